"""Time the covariance eigendecomposition of a CMA generation (torch.linalg.eigh = cuSOLVER) in fp64 and fp32."""
import json, torch
res = []
for n in (1024, 2048, 4096):
    A = torch.randn(n, n, device='cuda', dtype=torch.float64); C = A @ A.T / n + torch.eye(n, device='cuda', dtype=torch.float64)
    for dt in (torch.float64, torch.float32):
        X = C.to(dt)
        torch.linalg.eigh(X); torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 3
        ev[0].record()
        for _ in range(reps): w, V = torch.linalg.eigh(X)
        ev[1].record(); torch.cuda.synchronize()
        err = float(((V * w) @ V.T - X).abs().max() / X.abs().max())
        res.append(dict(n=n, dtype=str(dt).split('.')[-1], ms=round(ev[0].elapsed_time(ev[1]) / reps, 2), recon_rel=err))
print(json.dumps(res))
