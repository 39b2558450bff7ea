"""examples/nes_session.c: the C ABI used from plain C99 (no Python, no torch in the loop).  CPU: it compiles against
include/des_b200.h, links libdes_b200.so and fails loudly without a device.  GPU: its generations equal NESEngine's."""
import os
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, 'distributedes_b200')


def build_example(tmp_path):
    from distributedes_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build_library()
    exe = str(tmp_path / 'nes_session')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I' + os.path.join(REPO, 'include'),
                    os.path.join(REPO, 'examples', 'nes_session.c'), '-L' + PKG, '-ldes_b200', '-Wl,-rpath,' + PKG, '-o', exe], check=True)
    return exe


def test_c_example_builds_and_fails_loudly_without_a_device(tmp_path):
    import torch
    exe = build_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip('GPU present: covered by the gpu test')
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and 'no CPU fallback' in r.stderr


def example_inputs(P, T=128, d0=24, A=4):
    """The example's LCG inputs, restated."""
    s = np.uint32(12345)

    def unif(n):
        nonlocal s
        out = np.empty(n, dtype=np.float32)
        for i in range(n):
            s = np.uint32((int(s) * 1664525 + 1013904223) & 0xFFFFFFFF)
            out[i] = np.float32(np.float32((int(s) >> 8) + 0.5) * np.float32(1.0 / 16777216.0) * np.float32(2.0) - np.float32(1.0))
        return out
    theta = (np.float32(0.1) * unif(P)).astype(np.float32)
    obs = (np.float32(1.5) * unif(T * d0)).astype(np.float32).reshape(T, d0)
    target = (np.float32(0.9) * unif(T * A)).astype(np.float32).reshape(T, A)
    return theta, obs, target


@pytest.mark.gpu
@pytest.mark.parametrize('precision,name', [(0, 'fp32'), (2, 'f16x3')])
def test_c_example_matches_the_python_engine(tmp_path, precision, name):
    from distributedes_b200.engine import NESEngine
    from distributedes_b200.model import param_count
    exe = build_example(tmp_path)
    gens, N, H = 3, 256, 64
    r = subprocess.run([exe, str(gens), str(N), str(H), str(precision)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    means = [float(l.split()[-1]) for l in lines[:gens]]
    chk = float(lines[gens].split()[-1])
    P = param_count(24, H, 4)
    theta, obs, target = example_inputs(P)
    eng = NESEngine(state_dim=24, hidden=H, action_dim=4, pop_size=N, theta0=theta, obs=obs, target=target, sigma=0.1,
                    learning_rate=0.1, weight_decay=0.005, clip=1.0, seed=7, precision=name, device='cuda:0')
    for g in range(gens):
        eng.generation()
        # same inputs bit for bit, same kernels: only the printed precision separates the two
        assert abs(float(eng.fitness_all.double().mean()) - means[g]) <= 2e-6 * abs(means[g])
    w = (np.arange(P) % 7 + 1).astype(np.float64)
    ref = float((eng.theta_numpy().astype(np.float64) * w).sum())
    assert abs(chk - ref) <= 1e-6 * max(1.0, abs(ref))
