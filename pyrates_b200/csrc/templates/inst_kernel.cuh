// inst_kernel.cuh — hand-written solver template of the *instance* execution model.
//
// One CUDA thread owns the complete state vector of one sweep instance (node set x parameter set):
// the state, the Runge-Kutta stages and the swept parameters live in registers for the whole launch,
// the generated vector field `prb_rhs` (emitted by pyrates_b200/emit_inst.py from the model's compute
// graph) is inlined into every stage, and the only HBM traffic inside the time loop is the decimated,
// coalesced observer store ([sample][observer][instance] layout -> consecutive threads write
// consecutive addresses).  State / parameter arrays are structure-of-arrays ([variable][instance]).
//
// Replaces the reference's Python time loops BaseBackend._solve_euler / _solve_heun
// (pyrates/backend/base/base_backend.py:712-769): same sampling rule (record BEFORE the update when
// step % store_step == 0), same integer step counter `t = step + t0` handed to every stage.
//
// Expects from the generated prologue:  typedef real; PRB_NS, PRB_NP, PRB_NOBS, PRB_NRINGS, PRB_SOLVER,
// PRB_BLOCK; struct PrbThread; prb_thread_init(); prb_rhs(); prb_record().

#ifndef PRB_UNROLL_NS
#define PRB_UNROLL_NS _Pragma("unroll")
#endif

template <int N>
__device__ __forceinline__ void prb_axpy(real (&o)[N], const real (&y)[N], const real a, const real (&k)[N]) {
    PRB_UNROLL_NS
    for (int j = 0; j < N; ++j) o[j] = y[j] + a * k[j];
}

#ifndef PRB_MIN_BLOCKS
#define PRB_MIN_BLOCKS 1
#endif

extern "C" __global__ void __launch_bounds__(PRB_BLOCK, PRB_MIN_BLOCKS)
prb_integrate(const __grid_constant__ prb_kernel_args A)
{
    const long long B = A.n_inst;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;

    PrbThread T;
    prb_thread_init(T, A, b);

    real* __restrict__ Y = (real*)A.y;
    real y[PRB_NS];
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) y[j] = Y[(long long)j * B + b];

    real* __restrict__ out = (real*)A.out;
    const real dt = (real)A.dt;
    const real hdt = (real)(A.dt / 2);
    const real dt6 = (real)(A.dt / 6);
    (void)hdt; (void)dt6;

    const long long store_step = A.store_step;
    long long until_store = (store_step - (A.step0 % store_step)) % store_step;
    long long sample = A.sample0;
    long long t = A.step0 + A.t0;

    for (long long i = 0; i < A.n_steps; ++i, ++t) {
        if (until_store == 0) {
            prb_record(y, out + sample * (long long)PRB_NOBS * B, b, B);
            ++sample;
            until_store = store_step;
        }
        --until_store;

        real k[PRB_NS];
#if PRB_SOLVER == PRB_SOLVER_EULER
        prb_rhs(t, y, k, T);
        prb_axpy(y, y, dt, k);
#elif PRB_SOLVER == PRB_SOLVER_HEUN_REF
        // reference `heun`: rhs buffer is aliased, so both summands are the 2nd evaluation
        real ys[PRB_NS];
        prb_rhs(t, y, k, T);
        prb_axpy(ys, y, dt, k);
        prb_rhs(t, ys, k, T);
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) y[j] = y[j] + hdt * (k[j] + k[j]);
#elif PRB_SOLVER == PRB_SOLVER_HEUN
        real ys[PRB_NS], k2[PRB_NS];
        prb_rhs(t, y, k, T);
        prb_axpy(ys, y, dt, k);
        prb_rhs(t, ys, k2, T);
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) y[j] = y[j] + hdt * (k[j] + k2[j]);
#elif PRB_SOLVER == PRB_SOLVER_RK4
        real ys[PRB_NS], acc[PRB_NS];
        prb_rhs(t, y, k, T);                       // k1
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) { acc[j] = k[j]; ys[j] = y[j] + hdt * k[j]; }
        prb_rhs(t, ys, k, T);                      // k2
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) { acc[j] = acc[j] + (real)2 * k[j]; ys[j] = y[j] + hdt * k[j]; }
        prb_rhs(t, ys, k, T);                      // k3
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) { acc[j] = acc[j] + (real)2 * k[j]; ys[j] = y[j] + dt * k[j]; }
        prb_rhs(t, ys, k, T);                      // k4
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) y[j] = y[j] + dt6 * (acc[j] + k[j]);
#else
#error "unknown PRB_SOLVER"
#endif
    }

    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) Y[(long long)j * B + b] = y[j];
    prb_thread_exit(T, A, b);
}

// Single rhs evaluation dy = f(t, y) for every instance (the callable handed out by get_run_func).
extern "C" __global__ void __launch_bounds__(PRB_BLOCK)
prb_rhs_eval(const __grid_constant__ prb_kernel_args A)
{
    const long long B = A.n_inst;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    PrbThread T;
    prb_thread_init(T, A, b);
    const real* __restrict__ Y = (const real*)A.y;
    real y[PRB_NS], k[PRB_NS];
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) y[j] = Y[(long long)j * B + b];
    prb_rhs(A.step0 + A.t0, y, k, T);
    real* __restrict__ out = (real*)A.out;
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) out[(long long)j * B + b] = k[j];
}
