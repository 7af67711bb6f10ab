"""B200 micro-benchmark: fp64 throughput of mma.sync.m8n8k4.f64 (DMMA) vs DFMA, to choose the fp64 sweep-GEMM inner loop."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyrates_b200 import runtime as rt
from pyrates_b200.emit_inst import read_template

SRC = read_template("prb_kernel_abi.cuh") + r'''
extern "C" __global__ void __launch_bounds__(512, 1) probe_dmma(const __grid_constant__ prb_kernel_args A) {
    double c[16][2];
    for (int f = 0; f < 16; ++f) { c[f][0] = threadIdx.x * 1e-9; c[f][1] = f * 1e-9; }
    double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
    for (long long it = 0; it < A.n_steps; ++it) {
#pragma unroll
        for (int f = 0; f < 16; ++f)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c[f][0]), "+d"(c[f][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int f = 0; f < 16; ++f) s += c[f][0] + c[f][1];
    ((double*)A.out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
extern "C" __global__ void __launch_bounds__(512, 1) probe_dfma(const __grid_constant__ prb_kernel_args A) {
    double c[32];
    for (int f = 0; f < 32; ++f) c[f] = threadIdx.x * 1e-9 + f;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    for (long long it = 0; it < A.n_steps; ++it) {
#pragma unroll
        for (int f = 0; f < 32; ++f) c[f] = fma(c[f], a, b);
    }
    double s = 0; for (int f = 0; f < 32; ++f) s += c[f];
    ((double*)A.out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
'''
mod = rt.compile_module(SRC, "dmma_probe.cu")
d_out = rt.DeviceBuffer(148 * 512 * 8)
d_y = rt.DeviceBuffer(64)
for name, flop_per_thread_iter in (("probe_dmma", 16 * 256 * 2 / 32), ("probe_dfma", 32 * 2)):
    a = rt.KernelArgs()
    iters = 20000
    a.n_steps, a.store_step, a.n_inst, a.y, a.out = iters, 1, 1, d_y.ptr, d_out.ptr
    mod.run(name, a, block=512, grid=148)
    rt.synchronize()
    e0, e1 = rt.Event(), rt.Event()
    e0.record(); mod.run(name, a, block=512, grid=148); e1.record(); e1.synchronize()
    ms = e0.elapsed_ms(e1)
    print(f"[dmma_probe] {name}: {ms:.3f} ms, {148 * 512 * iters * flop_per_thread_iter / ms / 1e9:.2f} TFLOP/s fp64")
