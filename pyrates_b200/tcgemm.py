"""Host side of the tcgen05 coupling GEMM (csrc/templates/tc_gemm.cuh): operand tiling and the 3xTF32 split.

The sweep-batched dense coupling ``R[i, b] = sum_j W[i, j] x[j, b]`` (numpy.dot in the reference's generated
vector field, pyrates/backend/base/base_funcs.py:141-142, once per grid_search instance) runs on the 5th-gen
tensor cores for float32 models.  Both operands are stored in HBM already in the shared-memory layout
``tcgen05.mma`` reads (K-major, no swizzle: core matrices of 8 rows x 4 fp32), one contiguous tile per
(row block, K slab of 32), so a single 1-D bulk copy stages a tile.
"""
from __future__ import annotations

import numpy as np

from .emit_inst import read_template

BK = 32          # fp32 elements per K slab
ROWS_A = 128     # rows of W per tile (= TMEM lanes, UMMA M)

__all__ = ["BK", "ROWS_A", "split_tf32", "tile_operand", "untile_operand", "probe_source", "idesc_tf32"]


def split_tf32(x: np.ndarray):
    """x (float32) -> (hi, lo): hi keeps the top 10 explicit mantissa bits (tf32-exact), lo = x - hi is exact in fp32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    hi = (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    return hi, (x - hi).astype(np.float32)


def tile_operand(m: np.ndarray, rows: int) -> np.ndarray:
    """[R, K] row-major -> tiles [R/rows][K/BK][BK/4][rows/8][8][4] (flat float32).  R and K are zero-padded to
    multiples of ``rows`` / BK.  Element (r, k) of a tile sits at ((k//4)*(rows//8) + r//8)*32 + (r%8)*4 + k%4."""
    m = np.asarray(m, dtype=np.float32)
    R, K = m.shape
    Rp, Kp = -(-R // rows) * rows, -(-K // BK) * BK
    if (Rp, Kp) != (R, K):
        p = np.zeros((Rp, Kp), dtype=np.float32)
        p[:R, :K] = m
        m = p
    t = m.reshape(Rp // rows, rows // 8, 8, Kp // BK, BK // 4, 4)        # (rb, rg, r, ks, kc, c)
    return np.ascontiguousarray(t.transpose(0, 3, 4, 1, 2, 5)).reshape(-1)


def untile_operand(t: np.ndarray, R: int, K: int, rows: int) -> np.ndarray:
    Rp, Kp = -(-R // rows) * rows, -(-K // BK) * BK
    a = np.asarray(t, dtype=np.float32).reshape(Rp // rows, Kp // BK, BK // 4, rows // 8, 8, 4)
    return np.ascontiguousarray(a.transpose(0, 3, 4, 1, 2, 5)).reshape(Rp, Kp)[:R, :K]


def idesc_tf32(M: int, N: int) -> int:
    """tcgen05 instruction descriptor: kind::tf32, fp32 accumulate, A and B K-major (cute/arch/mma_sm100_desc.hpp)."""
    return (1 << 4) | (2 << 7) | (2 << 10) | ((N >> 3) << 17) | ((M >> 4) << 24)


def probe_source(bt: int = 128, nstage: int = 3) -> str:
    return "\n".join([read_template("prb_kernel_abi.cuh"), f"#define PRB_TC_BT {bt}", f"#define PRB_TC_NSTAGE {nstage}",
                      read_template("tc_gemm.cuh"), read_template("tc_probe.cuh")]) + "\n"
