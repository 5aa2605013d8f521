"""The kernels' FFT source (csrc/btkb_fft.cuh + csrc/btkb_f2.cuh) executed on the CPU: tests/host/fft_packed_host.cc runs the M/8
"threads" of a transform pair as std::threads over plain-array "shared memory" with a std::barrier, once through the scalar code
path the B200 measurements were made with (PK = false) and once through the packed 2 x fp32 path (PK = true: FADD2 / FMUL2 / FFMA2 on
the device, the same per-component IEEE operations as scalar code on the host).

Checked: (1) the scalar path is an M-point DFT (both directions, every M the pipeline accepts) to float precision against numpy in
fp64; (2) the packed path — the reformulated radix-4 / radix-8 butterflies (multiplications by +-i folded into the additions), the
two-instruction complex product, and the "base + compile-time offset" forms of the three shared-memory layouts — reproduces the
scalar path BIT FOR BIT; (3) every packed primitive equals the scalar expression it replaces on 20 000 random operands, and the
packed polyphase MAC equals the scalar one exactly; (4) the NLMS recurrence of the per-bin kernel (csrc/btkb_nlms_math.cuh: Yc = v^H x,
u.x, the projector step and the leaky update, C = 2, 4, 8) carried over 40 frames, and the Zelinski post-filter's cross-spectral
density recursions over 30 frames, give bit-identical states in both forms.
Test infrastructure only: the product has no CPU path."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    exe = str(tmp_path_factory.mktemp("fft") / "fft_packed_host")
    subprocess.check_call(["g++", "-O2", "-std=c++20", "-pthread", "-ffp-contract=off", "-I" + CUDA_INC,
                           os.path.join(ROOT, "tests", "host", "fft_packed_host.cc"), "-o", exe])
    return exe


@pytest.mark.parametrize("M", [256, 512, 1024, 2048])
@pytest.mark.parametrize("sign", [+1, -1])
@pytest.mark.parametrize("seed", [0, 1, 2, 1000])      # 0, 1: silent frames (signed zeros); 2: a single sample; otherwise Gaussian noise
def test_device_fft_source_on_the_cpu_scalar_and_packed(harness, M, sign, seed):
    out = subprocess.run([harness, str(M), str(sign), str(seed + (M + sign if seed >= 3 else 0))], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    rows = np.array([[float.fromhex(v) for v in ln.split()] for ln in out[:2 * M]])
    x = [rows[:M, 0] + 1j * rows[:M, 1], rows[:M, 2] + 1j * rows[:M, 3]]
    S = rows[M:]
    for i in range(2):
        # SIGN = +1: gsl_fft_complex_radix2_backward (unnormalised e^{+2 pi i nk/M}, modulated.cc:396); -1: ..._forward (modulated.cc:559)
        ref = np.fft.ifft(x[i]) * M if sign > 0 else np.fft.fft(x[i])
        got = S[:, 2 * i] + 1j * S[:, 2 * i + 1]
        assert np.linalg.norm(got - ref) <= 3e-7 * np.linalg.norm(ref)
    assert np.array_equal(S[:, :4], S[:, 4:]) and np.array_equal(np.signbit(S[:, :4]), np.signbit(S[:, 4:]))   # packed == scalar, bit for bit (hex floats), zeros with their signs
    tail = out[-1].split()
    assert tail[0] == "fold" and float.fromhex(tail[1]) == 0.0 and tail[2] == "primitives" and int(tail[3]) == 0
    assert tail[4] == "regs_differ" and int(tail[5]) == 0      # after the last pass every thread still holds its eight outputs in registers
    assert tail[6] == "nlms" and int(tail[7]) == 0             # 40 NLMS adaptation steps, 30 Zelinski CSD frames and 40 RLS steps (u and Pt; C = 2, 4, 8), scalar vs packed: identical states


@pytest.mark.parametrize("C", [2, 4, 8])
@pytest.mark.parametrize("reg,load", [(0.0, 1.0e6), (1.0e-2, 1.0e6), (1.0, 1.0e4)])   # the last pair makes the regularisation term matter
def test_device_rls_step_source_on_the_cpu_against_fp64(tmp_path, C, reg, load):
    """The RLS adaptation step k_perbin_rls compiles (csrc/btkb_nlms_math.cuh rls_core_step), run on the CPU in fp32 over 120 frames of one
    chain, against an fp64 evaluation of the same projector-form recursion (oracle/restate.py gsc_rls_projector, lib/pybeamformer.py:817-901,
    constraint_option = 0): beamformer output, a-posteriori error and active weights agree to fp32 accuracy; packed == scalar bit for bit."""
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    exe = str(tmp_path / "rls_host")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I" + CUDA_INC, os.path.join(ROOT, "tests", "host", "rls_host.cc"), "-o", exe])
    rng = np.random.default_rng(100 + C)
    T, mu, gamma = 120, 0.97, 0.04
    ph = rng.uniform(-3, 3, C)
    v = (np.exp(-1j * ph) / C).astype(np.complex64)                      # an array manifold vector, |v_c| = 1/C
    s = 3000.0 * (rng.standard_normal(T) + 1j * rng.standard_normal(T))   # look-direction source + interferer + sensor noise
    q = 1500.0 * (rng.standard_normal(T) + 1j * rng.standard_normal(T))
    vi = np.exp(-1j * rng.uniform(-3, 3, C))
    X = (s[:, None] * (C * v.astype(np.complex128))[None] + q[:, None] * vi[None] + 100.0 * (rng.standard_normal((T, C)) + 1j * rng.standard_normal((T, C)))).astype(np.complex64)
    txt = "\n".join("%.9g %.9g" % (z.real, z.imag) for z in np.concatenate([v, X.reshape(-1)]))
    out = subprocess.run([exe, str(C), str(T), repr(mu), repr(gamma), repr(reg), repr(load)], input=txt, capture_output=True, text=True, check=True).stdout.strip().split("\n")
    assert out[-1].split() == ["packed_mismatches", "0"]
    got = np.array([[float.fromhex(t) for t in ln.split()] for ln in out[:T]])
    # fp64 restatement of one chain (the formulas of restate.gsc_rls_projector)
    v64, X64 = v.astype(np.complex128), X.astype(np.complex128)
    def fp64(reg):
        Pt = (np.eye(C) - C * np.outer(v64, np.conj(v64))) / load
        u = np.zeros(C, np.complex128)
        ref = np.zeros((T, 2 + C), np.complex128)
        for t in range(T):
            x = X64[t]
            Yc = np.vdot(v64, x)
            xt = x - C * Yc * v64
            pv = Pt @ xt
            ip = np.real(np.vdot(xt, pv))
            g = pv / (mu + ip)
            Pn = (Pt - np.outer(g, np.conj(pv))) / mu
            ep = Yc - u @ x
            un = u + gamma * ep * np.conj(g)
            if reg > 0:
                un = un - reg * (np.conj(Pn) @ u)
            Pt, u = Pn, un
            ref[t, 0], ref[t, 1], ref[t, 2:] = Yc, Yc - u @ x, u
        return ref
    ref = fp64(reg)
    gotc = got[:, 0::2] + 1j * got[:, 1::2]
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel(gotc[:, 0], ref[:, 0]) < 1e-6          # Yc = v^H x
    assert rel(gotc[:, 1], ref[:, 1]) < 2e-5          # output of the canceller
    assert rel(gotc[-1, 2:], ref[-1, 2:]) < 5e-5      # active weights after 120 updates (fp32 precision-matrix recursion)
    assert np.linalg.norm(ref[-1, 2:]) > 0
    if reg >= 1.0 and C >= 4:
        assert rel(fp64(0.0)[-1, 2:], ref[-1, 2:]) > 2e-4   # ... and here the regularisation term is ten times the agreement asked for above
    print("rls host vs fp64: Yc %.1e  out %.1e  u %.1e  |u| %.2e" % (rel(gotc[:, 0], ref[:, 0]), rel(gotc[:, 1], ref[:, 1]), rel(gotc[-1, 2:], ref[-1, 2:]), np.linalg.norm(ref[-1, 2:])))
