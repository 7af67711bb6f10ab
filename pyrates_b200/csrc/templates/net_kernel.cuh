// net_kernel.cuh — hand-written template of the *network* execution model.
//
// One persistent cooperative grid (one CTA per SM, PRB_BLOCK threads) integrates a whole network whose
// state is too large for one thread: populations of N units with dense / pair-space / gather coupling.
// The time loop, all Runge-Kutta stages and the observer store run inside ONE launch; CTAs meet at a
// grid-wide barrier after every rhs evaluation (the coupling needs the complete stage vector) and at the
// few extra points where the generated code consumes a vector at another index than it was produced.
//
// Replaces BaseBackend._solve_euler/_solve_heun (pyrates/backend/base/base_backend.py:712-769) and the
// NumPy kernels behind the generated vector field: numpy.dot (base_funcs.py:141-142), einsum-wsum
// (:85-88), numpy.sum (:139), numpy.roll ring buffers (:143; pyrates/ir/circuit.py:400-425).
//
// Memory: the state lives in HBM/L2 as flat vectors (Y = y_n, S1..S3 = stage inputs, ACC = RK
// accumulator, YN = y_{n+1}); each element is owned by exactly one thread/warp per evaluation, so only
// the stage vectors are ever read by other CTAs.  W is streamed once per evaluation with 128-bit
// non-allocating loads; the source vector stays L1/L2 resident.
//
// Expects from the generated prologue: real, PRB_NS, PRB_NOBS, PRB_SOLVER, PRB_BLOCK, PRB_NRINGS,
// PRB_WORK_USER (element offset of the solver buffers inside `work`), prb_net_init_heads(),
// prb_net_advance_heads(), prb_eval<STAGE>() and PRB_OBS_OFF.

struct PrbNet {
    const real* yin;             // stage input of the current evaluation (written in-kernel: plain loads only)
    real* work;
    real* rings;
    const real* __restrict__ consts;
    const int* __restrict__ iconsts;
    const real* __restrict__ tables;
    real *Y, *YN, *S1, *S2, *S3, *ACC;
    real dt, hdt, dt6;
    long long t;
    int head[PRB_NRINGS > 0 ? PRB_NRINGS : 1];
    unsigned long long* bar;
    unsigned long long bar_target;
    long long gtid, gsize, gwarp, nwarps;
    int lane;
};

// ---- grid-wide barrier: monotonic counter, one arrival per CTA (host zeroes the counter before launch)
__device__ __forceinline__ void prb_grid_barrier(PrbNet& C) {
    __syncthreads();
    if (threadIdx.x == 0) {
        C.bar_target += gridDim.x;
        __threadfence();
        atomicAdd(C.bar, 1ULL);
        unsigned long long v;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(C.bar) : "memory");
        } while (v < C.bar_target);
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ real prb_warp_sum(real v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, result broadcast to every thread (used for full reductions computed redundantly per CTA)
__device__ __forceinline__ real prb_block_sum(real v) {
    __shared__ real red[32];
    __shared__ real total;
    v = prb_warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        real s = (l < (int)(blockDim.x >> 5)) ? red[l] : (real)0;
        s = prb_warp_sum(s);
        if (l == 0) total = s;
    }
    __syncthreads();
    return total;
}

__device__ __forceinline__ long long prb_ring_pos(int head, int col, int L) {
    int p = col - head;
    p += (p < 0) ? L : 0;
    p += (p < 0) ? L : 0;
    return (long long)p;
}

__device__ __forceinline__ real prb_max(real a, real b) { return (a != a || b != b) ? a + b : (a > b ? a : b); }
__device__ __forceinline__ real prb_min(real a, real b) { return (a != a || b != b) ? a + b : (a < b ? a : b); }
__device__ __forceinline__ real prb_sign(real a) { return (real)((a > (real)0) - (a < (real)0)); }

// ---- streaming row dot product: sum_j w[j] * x[j], lanes stride the row with 128-bit loads.
// w is read-only and touched once per evaluation -> non-coherent, no L1 allocation; x was written by other
// CTAs in the previous evaluation -> coherent loads (L1 was invalidated by the barrier's fence).
__device__ __forceinline__ double2 prb_ld_stream(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 prb_ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ double prb_row_dot(const double* __restrict__ w, const double* x, int m, int lane) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if ((m & 1) == 0 && ((((unsigned long long)w) | ((unsigned long long)x)) & 15ULL) == 0) {
        const double2* w2 = (const double2*)w;
        const double2* x2 = (const double2*)x;
        const int m2 = m >> 1;
        int j = lane;
        for (; j + 96 < m2; j += 128) {
            const double2 wa = prb_ld_stream(w2 + j), wb = prb_ld_stream(w2 + j + 32);
            const double2 wc = prb_ld_stream(w2 + j + 64), wd = prb_ld_stream(w2 + j + 96);
            const double2 xa = x2[j], xb = x2[j + 32], xc = x2[j + 64], xd = x2[j + 96];
            a0 = fma(wa.x, xa.x, a0); a0 = fma(wa.y, xa.y, a0);
            a1 = fma(wb.x, xb.x, a1); a1 = fma(wb.y, xb.y, a1);
            a2 = fma(wc.x, xc.x, a2); a2 = fma(wc.y, xc.y, a2);
            a3 = fma(wd.x, xd.x, a3); a3 = fma(wd.y, xd.y, a3);
        }
        for (; j < m2; j += 32) {
            const double2 wa = prb_ld_stream(w2 + j);
            const double2 xa = x2[j];
            a0 = fma(wa.x, xa.x, a0); a0 = fma(wa.y, xa.y, a0);
        }
    } else {
        for (int j = lane; j < m; j += 32) a0 = fma(__ldg(w + j), x[j], a0);
    }
    return prb_warp_sum((a0 + a1) + (a2 + a3));
}

__device__ __forceinline__ float prb_row_dot(const float* __restrict__ w, const float* x, int m, int lane) {
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if ((m & 3) == 0 && ((((unsigned long long)w) | ((unsigned long long)x)) & 15ULL) == 0) {
        const float4* w4 = (const float4*)w;
        const float4* x4 = (const float4*)x;
        const int m4 = m >> 2;
        int j = lane;
        for (; j + 32 < m4; j += 64) {
            const float4 wa = prb_ld_stream(w4 + j), wb = prb_ld_stream(w4 + j + 32);
            const float4 xa = x4[j], xb = x4[j + 32];
            a0 = fmaf(wa.x, xa.x, a0); a1 = fmaf(wa.y, xa.y, a1); a2 = fmaf(wa.z, xa.z, a2); a3 = fmaf(wa.w, xa.w, a3);
            a0 = fmaf(wb.x, xb.x, a0); a1 = fmaf(wb.y, xb.y, a1); a2 = fmaf(wb.z, xb.z, a2); a3 = fmaf(wb.w, xb.w, a3);
        }
        for (; j < m4; j += 32) {
            const float4 wa = prb_ld_stream(w4 + j);
            const float4 xa = x4[j];
            a0 = fmaf(wa.x, xa.x, a0); a1 = fmaf(wa.y, xa.y, a1); a2 = fmaf(wa.z, xa.z, a2); a3 = fmaf(wa.w, xa.w, a3);
        }
    } else {
        for (int j = lane; j < m; j += 32) a0 = fmaf(__ldg(w + j), x[j], a0);
    }
    return prb_warp_sum((a0 + a1) + (a2 + a3));
}

// ---- solver stage logic: what happens to the derivative k of state element e in stage STAGE.
// HEUN_REF reproduces the reference's aliased `heun` (y += dt/2*(k2+k2), base_backend.py:763-765).
template <int STAGE>
__device__ __forceinline__ void prb_emit_k(const PrbNet& C, const long long e, const real k) {
#if PRB_SOLVER == PRB_SOLVER_EULER
    C.YN[e] = C.Y[e] + C.dt * k;
#elif PRB_SOLVER == PRB_SOLVER_HEUN_REF
    if (STAGE == 0) C.S1[e] = C.Y[e] + C.dt * k;
    else            C.YN[e] = C.Y[e] + C.hdt * (k + k);
#elif PRB_SOLVER == PRB_SOLVER_HEUN
    if (STAGE == 0) { C.ACC[e] = k; C.S1[e] = C.Y[e] + C.dt * k; }
    else            C.YN[e] = C.Y[e] + C.hdt * (C.ACC[e] + k);
#elif PRB_SOLVER == PRB_SOLVER_RK4
    if (STAGE == 0)      { C.ACC[e] = k;                           C.S1[e] = C.Y[e] + C.hdt * k; }
    else if (STAGE == 1) { C.ACC[e] = C.ACC[e] + (real)2 * k;      C.S2[e] = C.Y[e] + C.hdt * k; }
    else if (STAGE == 2) { C.ACC[e] = C.ACC[e] + (real)2 * k;      C.S3[e] = C.Y[e] + C.dt * k; }
    else                 C.YN[e] = C.Y[e] + C.dt6 * (C.ACC[e] + k);
#endif
}

template <int STAGE> __device__ void prb_eval(PrbNet& C);     // generated

template <int STAGE>
__device__ __forceinline__ void prb_stage(PrbNet& C, const real* yin) {
    C.yin = yin;
    prb_eval<STAGE>(C);
    prb_net_advance_heads(C);
    prb_grid_barrier(C);
}

extern "C" __global__ void __launch_bounds__(PRB_BLOCK, 1)
prb_integrate_net(const __grid_constant__ prb_kernel_args A)
{
    PrbNet C;
    C.work = (real*)A.work;
    C.rings = (real*)A.rings;
    C.consts = (const real*)A.consts;
    C.iconsts = A.iconsts;
    C.tables = (const real*)A.tables;
    real* sol = C.work + (long long)PRB_WORK_USER;
    C.Y = sol; C.YN = sol + (long long)PRB_NS; C.S1 = sol + 2LL * PRB_NS; C.S2 = sol + 3LL * PRB_NS;
    C.S3 = sol + 4LL * PRB_NS; C.ACC = sol + 5LL * PRB_NS;
    C.dt = (real)A.dt; C.hdt = (real)(A.dt / 2); C.dt6 = (real)(A.dt / 6);
    C.bar = (unsigned long long*)A.sync;
    C.bar_target = 0;
    C.gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    C.gsize = (long long)gridDim.x * blockDim.x;
    C.lane = threadIdx.x & 31;
    C.gwarp = C.gtid >> 5;
    C.nwarps = C.gsize >> 5;
    prb_net_init_heads(C, A.eval0);

    real* extY = (real*)A.y;
    for (long long e = C.gtid; e < PRB_NS; e += C.gsize) C.Y[e] = extY[e];
    prb_grid_barrier(C);

    real* out = (real*)A.out;
    const long long store_step = A.store_step;
    long long until_store = (store_step - (A.step0 % store_step)) % store_step;
    long long sample = A.sample0;
    C.t = A.step0 + A.t0;

    for (long long i = 0; i < A.n_steps; ++i, ++C.t) {
        if (until_store == 0) {
            for (long long o = C.gtid; o < PRB_NOBS; o += C.gsize)
                out[sample * (long long)PRB_NOBS + o] = C.Y[PRB_OBS_INDEX(o)];
            ++sample;
            until_store = store_step;
        }
        --until_store;
#if PRB_SOLVER == PRB_SOLVER_EULER
        prb_stage<0>(C, C.Y);
#elif PRB_SOLVER == PRB_SOLVER_HEUN_REF || PRB_SOLVER == PRB_SOLVER_HEUN
        prb_stage<0>(C, C.Y);
        prb_stage<1>(C, C.S1);
#elif PRB_SOLVER == PRB_SOLVER_RK4
        prb_stage<0>(C, C.Y);
        prb_stage<1>(C, C.S1);
        prb_stage<2>(C, C.S2);
        prb_stage<3>(C, C.S3);
#endif
        real* tmp = C.Y; C.Y = C.YN; C.YN = tmp;
    }
    for (long long e = C.gtid; e < PRB_NS; e += C.gsize) extY[e] = C.Y[e];
}
