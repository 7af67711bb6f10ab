"""B200 check of the tcgen05 building blocks (tc_gemm.cuh) against NumPy: D = A . B^T for one 128 x BT tile per CTA.

Usage (GPU box): python scripts/tc_probe.py [--sweep]     --sweep also tries alternative descriptor encodings.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyrates_b200 import runtime as rt, tcgemm  # noqa: E402


def run(mod, A_t, B_t, MB, BT, KS, cfg, smem, reps=0):
    d_a = rt.DeviceBuffer.from_host(A_t)
    d_b = rt.DeviceBuffer.from_host(B_t)
    d_out = rt.DeviceBuffer(MB * 128 * BT * 4)
    d_out.zero()
    d_sync = rt.DeviceBuffer(64)
    d_sync.zero()
    d_ic = rt.DeviceBuffer.from_host(np.asarray(cfg, dtype=np.int32))
    d_y = rt.DeviceBuffer(64)
    a = rt.KernelArgs()
    a.n_steps, a.store_step, a.n_inst = 0, 1, 1
    a.y, a.out, a.consts, a.work, a.iconsts, a.sync = d_y.ptr, d_out.ptr, d_a.ptr, d_b.ptr, d_ic.ptr, d_sync.ptr
    mod.run("prb_tc_probe", a, block=192, grid=MB, smem=smem)
    rt.synchronize()
    if reps:
        e0, e1 = rt.Event(), rt.Event()
        e0.record()
        for _ in range(reps):
            mod.run("prb_tc_probe", a, block=192, grid=MB, smem=smem)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_ms(e1) / reps
        terms = cfg[6]
        fl = 2.0 * MB * 128 * BT * KS * 32 * terms
        byt = (MB * 128 + BT) * KS * 32 * 4 * 2
        print(f"[tc_probe] timing MB={MB} BT={BT} K={KS*32} terms={terms}: {ms*1e3:.1f} us/launch, "
              f"{fl/ms/1e9:.1f} TFLOP/s tf32 ({fl/ms/1e9/terms:.1f} effective), operand HBM {byt/ms/1e6:.0f} GB/s, "
              f"smem fill {MB*(128+BT)*KS*32*8/ms/1e6:.0f} GB/s")
    out = np.empty((MB * 128, BT), dtype=np.float32)
    d_out.download(out)
    sync = np.zeros(8, dtype=np.uint64)
    d_sync.download(sync)
    for b in (d_a, d_b, d_out, d_sync, d_ic, d_y):
        b.free()
    return out, int(sync[2])


def main():
    BT, NST = 128, 3
    MB, KS = 2, 8
    K = KS * tcgemm.BK
    rng = np.random.default_rng(0)
    A = rng.normal(size=(MB * 128, K)).astype(np.float32)
    B = rng.normal(size=(BT, K)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    a_hi, a_lo = tcgemm.split_tf32(A)
    b_hi, b_lo = tcgemm.split_tf32(B)
    A_t = np.concatenate([tcgemm.tile_operand(a_hi, 128), tcgemm.tile_operand(a_lo, 128)])
    B_t = np.concatenate([tcgemm.tile_operand(b_hi, BT), tcgemm.tile_operand(b_lo, BT)])
    mod = rt.compile_module(tcgemm.probe_source(BT, NST), "tc_probe.cu")
    smem = NST * (2 * 128 * 32 * 4 + 2 * BT * 32 * 4)
    idesc = tcgemm.idesc_tf32(128, BT)
    lboA, sboA, lboB, sboB = 128 * 16, 128, BT * 16, 128
    variants = [("canonical 1xTF32", [KS, lboA, sboA, lboB, sboB, idesc, 1, 2 * lboA, 2 * lboB]),
                ("canonical 3xTF32", [KS, lboA, sboA, lboB, sboB, idesc, 3, 2 * lboA, 2 * lboB])]
    scale = np.abs(ref).max()
    ok = True
    for name, cfg in variants:
        out, err = run(mod, A_t, B_t, MB, BT, KS, cfg, smem)
        e = np.abs(out - ref).max() / scale
        hi_ref = a_hi.astype(np.float64) @ b_hi.astype(np.float64).T
        e_hi = np.abs(out - hi_ref).max() / scale
        print(f"[tc_probe] {name}: err_flag={err} max scaled err vs fp64 = {e:.3e}, vs hi.hi = {e_hi:.3e}; "
              f"out[0,:4]={out[0, :4]} ref[0,:4]={ref[0, :4]}")
        if name == "canonical 3xTF32":
            ok = ok and err == 0 and e < 5e-6
        if name == "canonical 1xTF32":
            ok = ok and err == 0 and e < 5e-3
    if "--big" in sys.argv:
        ok = big() and ok
    print("[tc_probe]", "PASS" if ok else "FAIL")
    return 0 if ok else 1


def big():
    """K = 8192: accumulation error of the TMEM fp32 accumulator (zero-mean and all-positive data) + pacing."""
    ok = True
    for BT, NST in ((128, 3), (256, 2)):
        MB, KS = 148, 256
        K = KS * tcgemm.BK
        mod = rt.compile_module(tcgemm.probe_source(BT, NST), "tc_probe.cu")
        smem = NST * (2 * 128 * 32 * 4 + 2 * BT * 32 * 4)
        idesc = tcgemm.idesc_tf32(128, BT)
        lboA, sboA, lboB, sboB = 128 * 16, 128, BT * 16, 128
        for kind in ("zero-mean", "positive"):
            rng = np.random.default_rng(1)
            A = rng.normal(size=(MB * 128, K)).astype(np.float32)
            B = rng.normal(size=(BT, K)).astype(np.float32)
            if kind == "positive":
                A, B = np.abs(A), np.abs(B)
            a_hi, a_lo = tcgemm.split_tf32(A)
            b_hi, b_lo = tcgemm.split_tf32(B)
            A_t = np.concatenate([tcgemm.tile_operand(a_hi, 128), tcgemm.tile_operand(a_lo, 128)])
            B_t = np.concatenate([tcgemm.tile_operand(b_hi, BT), tcgemm.tile_operand(b_lo, BT)])
            sub = slice(0, 256)
            ref = A[sub].astype(np.float64) @ B.astype(np.float64).T
            f32 = A[sub] @ B.T
            for terms in (3, 1):
                cfg = [KS, lboA, sboA, lboB, sboB, idesc, terms, 2 * lboA, 2 * lboB]
                out, err = run(mod, A_t, B_t, MB, BT, KS, cfg, smem, reps=5 if kind == "zero-mean" else 0)
                d = out[sub] - ref
                print(f"[tc_probe] K={K} BT={BT} {kind} terms={terms}: err_flag={err} max scaled err {np.abs(d).max()/np.abs(ref).max():.3e} "
                      f"mean signed rel err {np.mean(d/ref) if kind=='positive' else float('nan'):.3e} "
                      f"(numpy fp32 matmul: {np.abs(f32-ref).max()/np.abs(ref).max():.3e})")
                ok = ok and err == 0
    return ok


if __name__ == "__main__":
    sys.exit(main())
