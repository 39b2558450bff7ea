mkdir -p gpurun_out
echo "== base (GEN_E1=2, both chunks at k=5)"; timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 600
for v in e1_0 e1_4 e1_4s e1_2s; do echo "== $v"; DES_LIB_PATH=distributedes_b200/libdes_b200_$v.so timeout 120 python scripts/time_eval.py 65536 256 f16x3 4 600; done
echo "== H=64 base / e1_0"; timeout 120 python scripts/time_eval.py 4096 64 f16x3 5 600; DES_LIB_PATH=distributedes_b200/libdes_b200_e1_0.so timeout 120 python scripts/time_eval.py 4096 64 f16x3 5 600
