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

// fast_math kernels read 2^(j/2048) from shared memory (divergent indices: a __constant__ read would serialise)
#ifdef PRB_USE_ETAB
#ifdef PRB_ETAB_FILE_SCOPE
#define PRB_DECL_ETAB
#else
#define PRB_DECL_ETAB __shared__ double prb_etab[PRB_ETAB_N + 2];
#endif
// entries N, N + 1 hold 2048/ln2: every thread reads ITS copy back after the barrier (T.ek0), which makes the multiplier a
// per-thread register value — as a compile-time constant it would sit in a uniform register and be copied into a register
// pair before every use (24 extra instructions per RK4 step of the JRC sweep)
#define PRB_STAGE_ETAB PRB_DECL_ETAB \
    for (int q = threadIdx.x; q < PRB_ETAB_N; q += blockDim.x) prb_etab[q] = PRB_EXP_TAB[q]; \
    if (threadIdx.x < 2) prb_etab[PRB_ETAB_N + threadIdx.x] = PRB_EXP_K[0]; \
    __syncthreads();
#if PRB_EXP_PINK0
#define PRB_SET_ETAB T.etab = prb_etab; T.etab_s = (unsigned int)__cvta_generic_to_shared(prb_etab); \
    T.ek0 = prb_etab[PRB_ETAB_N + (threadIdx.x & 1)];
#else
#define PRB_SET_ETAB T.etab = prb_etab; T.etab_s = (unsigned int)__cvta_generic_to_shared(prb_etab); T.ek0 = PRB_EXP_K[0];
#endif
#else
#define PRB_STAGE_ETAB
#define PRB_SET_ETAB
#endif

#ifdef PRB_EVAL_ONLY
// Jacobian programs (get_jacobian_func): n*n outputs from n state variables, evaluated once per call — no integrator.
#elif PRB_SOLVER == PRB_SOLVER_RK45
// ---- adaptive solver: the reference's solver='scipy' == scipy.integrate.solve_ivp(method='RK45', first_step=dt,
// t_eval=times) (BaseBackend._solve_scipy, pyrates/backend/base/base_backend.py:802-812).  Dormand-Prince 5(4) with
// SciPy's step controller (safety 0.9, factors in [0.2, 10], RMS error norm against atol + rtol*max(|y|,|y_new|)) and
// its 4th-order dense output evaluated at the sample times (scipy/integrate/_ivp/rk.py: rk_step, _step_impl,
// RkDenseOutput; ivp.py: t_eval loop).  One thread integrates one instance with its OWN step sequence; the seven
// stage vectors K live in registers.  Reference-specific: the generated rhs returns its in-place buffer and SciPy keeps
// f by reference, so K[0] of every attempt is the most recent rhs evaluation (after a rejected attempt:
// f(t + h_rej, y_rej)) — reproduced here because `f` below is exactly that buffer.  The state is integrated in double
// whatever the model precision (scipy/integrate/_ivp/base.py: check_arguments).
// Options (A.consts, doubles): [0] rtol [1] atol [2] t_start [3] t_bound [4] first_step [5] sample spacing
// [6] max_step [7] max attempts.  A.n_steps = number of samples; per-instance status (0 ok, 1 step too small,
// 2 attempt budget exhausted) goes to ((unsigned*)A.sync)[b].
extern "C" __global__ void __launch_bounds__(PRB_BLOCK, PRB_MIN_BLOCKS)
prb_integrate(const __grid_constant__ prb_kernel_args A)
{
    constexpr int NST = PRB_RK_NST, NP = PRB_RK_NP;      // stages (RK45: 6, RK23: 3, DOP853: 12) and dense-output terms (4 / 3 / 7)
#ifdef PRB_RK_DOP853
    // SciPy's DOP853 (pyrates_b200/dop853.py): 12 stages, error estimate from the embedded 5th- and 3rd-order formulas,
    // 7th-order dense output that needs three EXTRA stages (rows 13..15 of the extended tableau) in every step that contains
    // sample times.  Row 12 of K is f at the step end.
    constexpr int NK = PRB_RK_NX;
    constexpr double C_[NK] = PRB_RK_C_INIT;
    constexpr double A_[NK][NK - 1] = PRB_RK_A_INIT;
    constexpr double B_[NST] = PRB_RK_B_INIT;
    constexpr double E3_[NST + 1] = PRB_RK_E3_INIT;
    constexpr double E5_[NST + 1] = PRB_RK_E5_INIT;
    constexpr double D_[NP - 3][NK] = PRB_RK_D_INIT;
#else
    constexpr int NK = NST + 1;
    constexpr double C_[NST] = PRB_RK_C_INIT;            // tableau emitted by pyrates_b200/rk45.py (SciPy's RK45 / RK23)
    constexpr double A_[NST][NST - 1] = PRB_RK_A_INIT;
    constexpr double B_[NST] = PRB_RK_B_INIT;
    constexpr double E_[NST + 1] = PRB_RK_E_INIT;
    constexpr double P_[NST + 1][NP] = PRB_RK_P_INIT;    // dense output: y(t_old + x h) = y_old + h * sum_c Q[:, c] x^(c+1), Q = K^T P
#endif
    const long long B = A.n_inst;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    PRB_STAGE_ETAB
    if (b >= B) return;
    PrbThread T;
    prb_thread_init(T, A, b);
    PRB_SET_ETAB
    const double* __restrict__ opt = (const double*)A.consts;
    const double rtol = opt[0], atol = opt[1], t_bound = opt[3], te_step = opt[5], max_step = opt[6];
    const long long max_attempts = (long long)opt[7];
    double t = opt[2], h_abs = opt[4];

    real* __restrict__ Y = (real*)A.y;
    double y[PRB_NS], f[PRB_NS];
    areal yr[PRB_NS], kr[PRB_NS];          // rhs input / output in the rhs arithmetic type (double for float32 models too,
    real yo[PRB_NS];                       // see emit_inst.py), recorded samples in the storage type
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) { y[j] = (double)Y[(long long)j * B + b]; f[j] = 0.0; }
    bool need_f = true;                    // f must be (re)computed as f(t, y): at the start (self.f = fun(t0, y0)) and after
                                           // a rejected attempt whose aliased f overflowed (see below); ONE inlined call site

    real* __restrict__ out = (real*)A.out;
    long long sample = A.sample0;
    const long long n_samples = A.sample0 + A.n_steps;
    long long attempts = 0;
    unsigned int status = 0;

    while (t != t_bound && status == 0) {
        const double min_step = 10.0 * fabs(nextafter(t, __longlong_as_double(0x7ff0000000000000LL)) - t);
        if (h_abs > max_step) h_abs = max_step;
        else if (h_abs < min_step) h_abs = min_step;
        bool rejected = false;
        double K[NK][PRB_NS], y_new[PRB_NS];
        double h = 0.0, t_new = t;
        for (;;) {
            if (need_f) {
                PRB_UNROLL_NS
                for (int j = 0; j < PRB_NS; ++j) yr[j] = (areal)y[j];
                prb_rhs((prb_time)t, yr, kr, T);
                PRB_UNROLL_NS
                for (int j = 0; j < PRB_NS; ++j) f[j] = (double)kr[j];
                need_f = false;
            }
            if (h_abs < min_step) { status = 1; break; }
            if (++attempts > max_attempts) { status = 2; break; }
            t_new = t + h_abs;
            if (t_new - t_bound > 0.0) t_new = t_bound;
            h = t_new - t;
            h_abs = fabs(h);
            PRB_UNROLL_NS
            for (int j = 0; j < PRB_NS; ++j) K[0][j] = f[j];
#pragma unroll
            for (int s = 1; s < NST; ++s) {
                PRB_UNROLL_NS
                for (int j = 0; j < PRB_NS; ++j) {
                    double d = 0.0;
#pragma unroll
                    for (int q = 0; q < s; ++q) d += K[q][j] * A_[s][q];
                    yr[j] = (areal)(y[j] + d * h);
                }
                prb_rhs((prb_time)(t + C_[s] * h), yr, kr, T);
                PRB_UNROLL_NS
                for (int j = 0; j < PRB_NS; ++j) K[s][j] = (double)kr[j];
            }
            PRB_UNROLL_NS
            for (int j = 0; j < PRB_NS; ++j) {
                double d = 0.0;
#pragma unroll
                for (int q = 0; q < NST; ++q) d += K[q][j] * B_[q];
                y_new[j] = y[j] + h * d;
                yr[j] = (areal)y_new[j];
            }
            prb_rhs((prb_time)(t + h), yr, kr, T);
#ifdef PRB_RK_DOP853
            double sq5 = 0.0, sq3 = 0.0;
            PRB_UNROLL_NS
            for (int j = 0; j < PRB_NS; ++j) {
                f[j] = (double)kr[j];                                  // the aliased buffer: newest evaluation
                K[NST][j] = f[j];
                double e5 = 0.0, e3 = 0.0;
#pragma unroll
                for (int q = 0; q <= NST; ++q) { e5 += K[q][j] * E5_[q]; e3 += K[q][j] * E3_[q]; }
                const double scale = atol + fmax(fabs(y[j]), fabs(y_new[j])) * rtol;
                e5 /= scale; e3 /= scale;
                sq5 += e5 * e5; sq3 += e3 * e3;
            }
            // DOP853._estimate_error_norm: |h| * ||err5||^2 / sqrt((||err5||^2 + 0.01 ||err3||^2) * n)
            const double error_norm = (sq5 == 0.0 && sq3 == 0.0) ? 0.0
                : fabs(h) * sq5 / sqrt((sq5 + 0.01 * sq3) * (double)PRB_NS);
#else
            double sq = 0.0;
            PRB_UNROLL_NS
            for (int j = 0; j < PRB_NS; ++j) {
                f[j] = (double)kr[j];                                  // the aliased buffer: newest evaluation
                K[NST][j] = f[j];
                double e = 0.0;
#pragma unroll
                for (int q = 0; q <= NST; ++q) e += K[q][j] * E_[q];
                const double scale = atol + fmax(fabs(y[j]), fabs(y_new[j])) * rtol;
                const double r = e * h / scale;
                sq += r * r;
            }
            const double error_norm = sqrt(sq) / sqrt((double)PRB_NS);
#endif
            if (error_norm < 1.0) {
                double factor = (error_norm == 0.0) ? 10.0 : fmin(10.0, 0.9 * pow(error_norm, PRB_RK_EXPONENT));
                if (rejected) factor = fmin(1.0, factor);
                h_abs *= factor;
                break;
            }
            h_abs *= fmax(0.2, 0.9 * pow(error_norm, PRB_RK_EXPONENT));
            rejected = true;
            // The aliased buffer now holds f(t + h_rej, y_rej).  If that attempt overflowed, the reference is stuck with a
            // NaN K[0] for ever and solve_ivp dies with "step size too small" (seen with its own test_3_2 inputs, on
            // either side of a rounding coin flip).  Whenever the reference survives this value is finite and used as is;
            // otherwise recover with the true f(t, y) instead of reproducing the crash.
            bool finite = true;
            PRB_UNROLL_NS
            for (int j = 0; j < PRB_NS; ++j) finite = finite && (fabs(f[j]) <= 1.7976931348623157e308);
            need_f = !finite;
        }
        if (status != 0) break;
#ifdef PRB_RK_DOP853
        // samples inside (t_old, t_new]: Dop853DenseOutput.  Three extra stages (each overwrites the reference's aliased rhs
        // buffer, so `f` ends up as the LAST of them — that is what the interpolant uses as "f at the step end" and what
        // the next attempt starts from), F[0..2] from the step's end points, F[3..6] = h * D . K, then the nested product
        // y = y_old + x (F0 + (1-x) (F1 + x (F2 + (1-x) (F3 + x (F4 + (1-x) (F5 + x F6))))))
        if (sample < n_samples && (double)(sample - A.sample0) * te_step <= t_new) {
#pragma unroll
            for (int s = NST + 1; s < NK; ++s) {
                PRB_UNROLL_NS
                for (int j = 0; j < PRB_NS; ++j) {
                    double d = 0.0;
#pragma unroll
                    for (int q = 0; q < s; ++q) d += K[q][j] * A_[s][q];
                    yr[j] = (areal)(y[j] + d * h);
                }
                prb_rhs((prb_time)(t + C_[s] * h), yr, kr, T);
                PRB_UNROLL_NS
                for (int j = 0; j < PRB_NS; ++j) { K[s][j] = (double)kr[j]; f[j] = K[s][j]; }
            }
            double F[NP][PRB_NS];
            PRB_UNROLL_NS
            for (int j = 0; j < PRB_NS; ++j) {
                const double dy_ = y_new[j] - y[j];
                F[0][j] = dy_;
                F[1][j] = h * K[0][j] - dy_;
                F[2][j] = 2.0 * dy_ - h * (f[j] + K[0][j]);
#pragma unroll
                for (int c = 0; c < NP - 3; ++c) {
                    double d = 0.0;
#pragma unroll
                    for (int q = 0; q < NK; ++q) d += D_[c][q] * K[q][j];
                    F[3 + c][j] = h * d;
                }
            }
            while (sample < n_samples && (double)(sample - A.sample0) * te_step <= t_new) {
                const double x = ((double)(sample - A.sample0) * te_step - t) / h;
                PRB_UNROLL_NS
                for (int j = 0; j < PRB_NS; ++j) {
                    double acc = 0.0;
#pragma unroll
                    for (int i = 0; i < NP; ++i) {
                        acc += F[NP - 1 - i][j];
                        acc *= (i % 2 == 0) ? x : (1.0 - x);
                    }
                    yo[j] = (real)(acc + y[j]);
                }
                prb_record(yo, out + sample * (long long)PRB_NOBS * B, b, B);
                ++sample;
            }
        }
#else
        // samples inside (t_old, t_new]: dense output  y_old + h * Q . (x, x^2, .., x^NP),  Q = K^T P
        while (sample < n_samples && (double)(sample - A.sample0) * te_step <= t_new) {
            const double x = ((double)(sample - A.sample0) * te_step - t) / h;
            double pw[NP];
            pw[0] = x;
#pragma unroll
            for (int c = 1; c < NP; ++c) pw[c] = pw[c - 1] * x;          // numpy.cumprod
            PRB_UNROLL_NS
            for (int j = 0; j < PRB_NS; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < NP; ++c) {
                    double qc = 0.0;
#pragma unroll
                    for (int s = 0; s <= NST; ++s) qc += K[s][j] * P_[s][c];
                    acc += qc * pw[c];                                   // numpy.dot(Q, p): left to right
                }
                yo[j] = (real)(h * acc + y[j]);
            }
            prb_record(yo, out + sample * (long long)PRB_NOBS * B, b, B);
            ++sample;
        }
#endif
        t = t_new;
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) y[j] = y_new[j];
    }
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) Y[(long long)j * B + b] = (real)y[j];
    if (A.sync) A.sync[b] = status | ((unsigned int)(attempts > 0xFFFFFFF ? 0xFFFFFFF : attempts) << 4);   // low 4 bits: status, rest: rhs attempts
    prb_thread_exit(T, A, b);
}
#elif PRB_SOLVER == PRB_SOLVER_DOPRI5
// ---- adaptive solver of delay-differential models: the reference's solver='scipy' with a state history ==
// BaseBackend._solve_scipy_dde (pyrates/backend/base/base_backend.py:771-800):
//     solver = scipy.integrate.ode(rhs).set_integrator('dopri5', first_step=dt, nsteps=50000); set_solout(hist.update)
//     for t_out in times: solver.integrate(t_out); state_rec[i] = solver.y
// i.e. Hairer's DOPRI5 core (rtol 1e-6, atol 1e-12, safety 0.9, step factors in [0.2, 10], beta 0.04 — the reference
// forwards none of the caller's tolerances), RESTARTED for every sample interval with h = first_step, and every accepted
// step (plus the starting point of every call) appended to the history the rhs interpolates in (prb_hist).  One thread
// integrates one instance; the state is double whatever the model precision (scipy.integrate.ode converts it).
// Options (A.consts, doubles): [0] rtol [1] atol [2] t_start [4] first_step [5] sample spacing [7] nsteps
// [8] history capacity [9..] float64 initial condition of the history.  A.work: history ring.  Status per instance
// (A.sync): 0 ok, 1 step size too small, 2 more than nsteps steps in one call, 3 history ring too small (host retries).
extern "C" __global__ void __launch_bounds__(PRB_BLOCK, PRB_MIN_BLOCKS)
prb_integrate(const __grid_constant__ prb_kernel_args A)
{
    constexpr double C2 = 0.2, C3 = 0.3, C4 = 0.8, C5 = 8.0 / 9.0;
    constexpr double A21 = 0.2, A31 = 3.0 / 40.0, A32 = 9.0 / 40.0, A41 = 44.0 / 45.0, A42 = -56.0 / 15.0, A43 = 32.0 / 9.0,
        A51 = 19372.0 / 6561.0, A52 = -25360.0 / 2187.0, A53 = 64448.0 / 6561.0, A54 = -212.0 / 729.0,
        A61 = 9017.0 / 3168.0, A62 = -355.0 / 33.0, A63 = 46732.0 / 5247.0, A64 = 49.0 / 176.0, A65 = -5103.0 / 18656.0,
        A71 = 35.0 / 384.0, A73 = 500.0 / 1113.0, A74 = 125.0 / 192.0, A75 = -2187.0 / 6784.0, A76 = 11.0 / 84.0;
    constexpr double E1 = 71.0 / 57600.0, E3 = -71.0 / 16695.0, E4 = 71.0 / 1920.0, E5 = -17253.0 / 339200.0,
        E6 = 22.0 / 525.0, E7 = -1.0 / 40.0;
    constexpr double SAFE = 0.9, BETA = 0.04, EXPO1 = 0.2 - BETA * 0.75, FACC1 = 1.0 / 0.2, FACC2 = 1.0 / 10.0,
        UROUND = 2.3e-16;
    const long long B = A.n_inst;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    PRB_STAGE_ETAB
    if (b >= B) return;
    PrbThread T;
    prb_thread_init(T, A, b);
    PRB_SET_ETAB
    const double* __restrict__ opt = (const double*)A.consts;
    const double rtol = opt[0], atol = opt[1], first_step = opt[4], te_step = opt[5];
    const long long nmax = (long long)opt[7];
    double x = opt[2];

    real* __restrict__ Y = (real*)A.y;
    double y[PRB_NS], y1[PRB_NS], k1[PRB_NS], k2[PRB_NS], k3[PRB_NS], k4[PRB_NS], k5[PRB_NS], k6[PRB_NS];
    areal yr[PRB_NS], kr[PRB_NS];
    real yo[PRB_NS];
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) y[j] = (double)Y[(long long)j * B + b];

    real* __restrict__ out = (real*)A.out;
    long long steps_total = 0;
    unsigned int status = 0;

#define PRB_D5_STAGE(tt, kout, expr)                                              \
    PRB_UNROLL_NS                                                                 \
    for (int j = 0; j < PRB_NS; ++j) yr[j] = (areal)(y[j] + (expr));              \
    prb_rhs((prb_time)(tt), yr, kr, T);                                           \
    PRB_UNROLL_NS                                                                 \
    for (int j = 0; j < PRB_NS; ++j) kout[j] = (double)kr[j];

    for (long long sample = 0; sample < A.n_steps && status == 0; ++sample) {
        const double xend = (double)sample * te_step;               // numpy.linspace(0, T, n, endpoint=False)[sample]
        // ---- one dopri5 call from x to xend (DOPCOR)
        double h = first_step, facold = 1e-4;
        const double posneg = (xend - x >= 0.0) ? 1.0 : -1.0, hmax = fabs(xend - x);
        bool last = false, reject = false;
        long long nstep = 0;
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) yr[j] = (areal)y[j];
        prb_rhs((prb_time)x, yr, kr, T);
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) k1[j] = (double)kr[j];
        prb_hist_push(T, x, y);                                      // solout at the start of the call
        for (;;) {
            if (nstep > nmax) { status = 2; break; }
            if (0.1 * fabs(h) <= fabs(x) * UROUND) { status = 1; break; }
            if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = true; }
            ++nstep;
            PRB_D5_STAGE(x + C2 * h, k2, h * A21 * k1[j])
            PRB_D5_STAGE(x + C3 * h, k3, h * (A31 * k1[j] + A32 * k2[j]))
            PRB_D5_STAGE(x + C4 * h, k4, h * (A41 * k1[j] + A42 * k2[j] + A43 * k3[j]))
            PRB_D5_STAGE(x + C5 * h, k5, h * (A51 * k1[j] + A52 * k2[j] + A53 * k3[j] + A54 * k4[j]))
            const double xph = x + h;
            PRB_D5_STAGE(xph, k6, h * (A61 * k1[j] + A62 * k2[j] + A63 * k3[j] + A64 * k4[j] + A65 * k5[j]))
            PRB_UNROLL_NS
            for (int j = 0; j < PRB_NS; ++j) {
                y1[j] = y[j] + h * (A71 * k1[j] + A73 * k3[j] + A74 * k4[j] + A75 * k5[j] + A76 * k6[j]);
                yr[j] = (areal)y1[j];
            }
            prb_rhs((prb_time)xph, yr, kr, T);                       // k7 (first stage of the next step)
            double err = 0.0;
            PRB_UNROLL_NS
            for (int j = 0; j < PRB_NS; ++j) {
                k2[j] = (double)kr[j];
                const double e = (E1 * k1[j] + E3 * k3[j] + E4 * k4[j] + E5 * k5[j] + E6 * k6[j] + E7 * k2[j]) * h;
                const double sk = atol + rtol * fmax(fabs(y[j]), fabs(y1[j]));
                err += (e / sk) * (e / sk);
            }
            err = sqrt(err / (double)PRB_NS);
            const double fac11 = pow(err, EXPO1);
            double fac = fac11 / pow(facold, BETA);
            fac = fmax(FACC2, fmin(FACC1, fac / SAFE));
            double hnew = h / fac;
            if (err <= 1.0) {
                facold = fmax(err, 1e-4);
                PRB_UNROLL_NS
                for (int j = 0; j < PRB_NS; ++j) { k1[j] = k2[j]; y[j] = y1[j]; }
                x = xph;
                prb_hist_push(T, x, y);                              // solout after every accepted step
                if (last) break;
                if (fabs(hnew) > hmax) hnew = posneg * hmax;
                if (reject) hnew = posneg * fmin(fabs(hnew), fabs(h));
                reject = false;
            } else {
                hnew = h / fmin(FACC1, fac11 / SAFE);
                reject = true;
                last = false;
            }
            h = hnew;
        }
        steps_total += nstep;
        if (T.hovf) status = 3;
        if (status != 0) break;
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) yo[j] = (real)y[j];
        prb_record(yo, out + (A.sample0 + sample) * (long long)PRB_NOBS * B, b, B);
    }
#undef PRB_D5_STAGE
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) Y[(long long)j * B + b] = (real)y[j];
    if (A.sync) A.sync[b] = status | ((unsigned int)(steps_total > 0xFFFFFFF ? 0xFFFFFFF : steps_total) << 4);
    prb_thread_exit(T, A, b);
}
#else
#ifdef PRB_Y_SMEM
__device__ __forceinline__ real prb_ysm_ld(const PrbThread& T, const int j) {
    real v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(T.ysm_s + (unsigned int)(j * PRB_BLOCK * 8)));
    return v;
}
__device__ __forceinline__ void prb_ysm_st(const PrbThread& T, const int j, const real v) {
    asm volatile("st.shared.f64 [%0], %1;" :: "r"(T.ysm_s + (unsigned int)(j * PRB_BLOCK * 8)), "d"(v) : "memory");
}
#endif
// One solver step y -> yn.  SPEC (fast_math kernels of pure vector fields, PRB_SPEC_STEP): every stage runs the speculative
// rhs, which only ORs range violations of its exp / reciprocal arguments into `bad`; the caller looks at the flag once per
// step and recomputes the step with the exact rhs (prb_step_exact) in the rare case it is set — the stages of a step stay one
// basic block instead of ending in a reconvergence point each.
template <bool SPEC>
__device__ __forceinline__ void prb_stage(const long long t, const real (&y)[PRB_NS], real (&k)[PRB_NS], PrbThread& T, PrbBad& bad) {
#ifdef PRB_SPEC_STEP
    if (SPEC) prb_rhs_spec(t, y, k, T, bad);
    else      prb_rhs_exact(t, y, k, T);
#else
    prb_rhs(t, y, k, T);
#endif
}

// FINAL = false: `yn` receives the increment direction d of the step instead of the new state (y_new = y + prb_step_coef * d);
// the caller finishes the update itself (PRB_SPEC_STEP: after it has looked at `bad`, so that both outcomes write y in place).
#if PRB_SOLVER == PRB_SOLVER_EULER
#define prb_step_coef(dt, hdt, dt6) (dt)
#elif PRB_SOLVER == PRB_SOLVER_RK4
#define prb_step_coef(dt, hdt, dt6) (dt6)
#else
#define prb_step_coef(dt, hdt, dt6) (hdt)
#endif
template <bool SPEC, bool FINAL = true>
__device__ __forceinline__ void prb_step(const long long t, const real (&y)[PRB_NS], real (&yn)[PRB_NS], PrbThread& T,
                                         const real dt, const real hdt, const real dt6, PrbBad& bad) {
    real k[PRB_NS];
    (void)hdt; (void)dt6;
#if PRB_SOLVER == PRB_SOLVER_EULER
    prb_stage<SPEC>(t, y, k, T, bad);
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) yn[j] = FINAL ? y[j] + dt * k[j] : k[j];
#elif PRB_SOLVER == PRB_SOLVER_HEUN_REF
    // reference `heun`: rhs buffer is aliased, so both summands are the 2nd evaluation
    real ys[PRB_NS];
    prb_stage<SPEC>(t, y, k, T, bad);
    prb_axpy(ys, y, dt, k);
    prb_stage<SPEC>(t, ys, k, T, bad);
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) yn[j] = FINAL ? y[j] + hdt * (k[j] + k[j]) : k[j] + k[j];
#elif PRB_SOLVER == PRB_SOLVER_HEUN
    real ys[PRB_NS], k2[PRB_NS];
    prb_stage<SPEC>(t, y, k, T, bad);
    prb_axpy(ys, y, dt, k);
    prb_stage<SPEC>(t, ys, k2, T, bad);
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) yn[j] = FINAL ? y[j] + hdt * (k[j] + k2[j]) : k[j] + k2[j];
#elif PRB_SOLVER == PRB_SOLVER_RK4 && defined(PRB_Y_SMEM)
    // Register-lean variant (opt-in, PRB_Y_SMEM): the base state y lives in shared memory ([state][thread], one 8-byte word
    // per thread and state variable: conflict-free) and is re-read where a stage needs it, so that only the stage input, the
    // accumulator and the fresh derivatives are live across an evaluation.  `y` / `yn` are unused here: the caller passes
    // the shared-memory column of this thread.
    real ys[PRB_NS], acc[PRB_NS];
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) ys[j] = prb_ysm_ld(T, j);
    prb_stage<SPEC>(t, ys, k, T, bad);                      // k1
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) { acc[j] = k[j]; ys[j] = prb_ysm_ld(T, j) + hdt * k[j]; }
    prb_stage<SPEC>(t, ys, k, T, bad);                      // k2
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) { acc[j] = acc[j] + (real)2 * k[j]; ys[j] = prb_ysm_ld(T, j) + hdt * k[j]; }
    prb_stage<SPEC>(t, ys, k, T, bad);                      // k3
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) { acc[j] = acc[j] + (real)2 * k[j]; ys[j] = prb_ysm_ld(T, j) + dt * k[j]; }
    prb_stage<SPEC>(t, ys, k, T, bad);                      // k4
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) yn[j] = prb_ysm_ld(T, j) + dt6 * (acc[j] + k[j]);
    (void)y;
#elif PRB_SOLVER == PRB_SOLVER_RK4 && defined(PRB_INTEG_PAIRS)
    if constexpr (FINAL) {                                  // the out-of-line exact step: plain RK4
        real ys[PRB_NS], acc[PRB_NS];
        prb_stage<SPEC>(t, y, k, T, bad);
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) { acc[j] = k[j]; ys[j] = y[j] + hdt * k[j]; }
        prb_stage<SPEC>(t, ys, k, T, bad);
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) { acc[j] = acc[j] + (real)2 * k[j]; ys[j] = y[j] + hdt * k[j]; }
        prb_stage<SPEC>(t, ys, k, T, bad);
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) { acc[j] = acc[j] + (real)2 * k[j]; ys[j] = y[j] + dt * k[j]; }
        prb_stage<SPEC>(t, ys, k, T, bad);
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) yn[j] = y[j] + dt6 * (acc[j] + k[j]);
    } else {
    // Integrator chains (PRB_INTEG[j] = b >= 0: dy_j == y_b, so k_j of a stage IS that stage's y_b).  Such a state needs no
    // accumulator:  y_j' = y_j + dt y_b + dt^2/6 (k1 + k2 + k3)_b  exactly, and every other state takes its increment from the
    // same sum:  k1 + 2 k2 + 2 k3 + k4 = 2 S + (k4 - k1),  S = k1 + k2 + k3.  Only instantiated with FINAL = false: `yn`
    // receives S of the SOURCE state for chain states (the in-place finish of the time loop applies the formula) and
    // 2 S + (k4 - k1) otherwise.
    real ys[PRB_NS], k1[PRB_NS], S[PRB_NS];
    prb_stage<SPEC>(t, y, k, T, bad);                       // k1
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) { k1[j] = k[j]; ys[j] = y[j] + hdt * k[j]; }
    prb_stage<SPEC>(t, ys, k, T, bad);                      // k2
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) { S[j] = k1[j] + k[j]; ys[j] = y[j] + hdt * k[j]; }
    prb_stage<SPEC>(t, ys, k, T, bad);                      // k3
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) { S[j] = S[j] + k[j]; ys[j] = y[j] + dt * k[j]; }
    prb_stage<SPEC>(t, ys, k, T, bad);                      // k4
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j)
        yn[j] = PRB_INTEG[j] >= 0 ? S[PRB_INTEG[j]] : (real)2 * S[j] + (k[j] - k1[j]);
    }
#elif PRB_SOLVER == PRB_SOLVER_RK4
    real ys[PRB_NS], acc[PRB_NS];
    prb_stage<SPEC>(t, y, k, T, bad);                       // k1
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) { acc[j] = k[j]; ys[j] = y[j] + hdt * k[j]; }
    prb_stage<SPEC>(t, ys, k, T, bad);                      // k2
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) { acc[j] = acc[j] + (real)2 * k[j]; ys[j] = y[j] + hdt * k[j]; }
    prb_stage<SPEC>(t, ys, k, T, bad);                      // k3
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) { acc[j] = acc[j] + (real)2 * k[j]; ys[j] = y[j] + dt * k[j]; }
    prb_stage<SPEC>(t, ys, k, T, bad);                      // k4
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) yn[j] = FINAL ? y[j] + dt6 * (acc[j] + k[j]) : acc[j] + k[j];
#else
#error "unknown PRB_SOLVER"
#endif
}

#ifdef PRB_SPEC_STEP
__device__ __noinline__ void prb_step_exact(const long long t, const real (&y)[PRB_NS], real (&yn)[PRB_NS], PrbThread& T,
                                            const real dt, const real hdt, const real dt6) {
    PrbBad unused;
    prb_step<false>(t, y, yn, T, dt, hdt, dt6, unused);
}
#endif

extern "C" __global__ void __launch_bounds__(PRB_BLOCK, PRB_MIN_BLOCKS)
prb_integrate(const __grid_constant__ prb_kernel_args A)
{
    const long long B = A.n_inst;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    PRB_STAGE_ETAB
    if (b >= B) return;

    PrbThread T;
    prb_thread_init(T, A, b);
    PRB_SET_ETAB

    real* __restrict__ Y = (real*)A.y;
    real y[PRB_NS];
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) y[j] = Y[(long long)j * B + b];
#ifdef PRB_Y_SMEM
    __shared__ real prb_ysm[PRB_NS][PRB_BLOCK];
    T.ysm_s = (unsigned int)__cvta_generic_to_shared(&prb_ysm[0][threadIdx.x]);
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) prb_ysm_st(T, j, y[j]);
#endif

    real* __restrict__ out = (real*)A.out;
    const real dt = (real)A.dt;
    const real hdt = (real)(A.dt / 2);
    const real dt6 = (real)(A.dt / 6);
    (void)hdt; (void)dt6;

    const long long store_step = A.store_step;
    long long until_store = (store_step - (A.step0 % store_step)) % store_step;
    long long sample = A.sample0;
    long long t = A.step0 + A.t0;

#ifdef PRB_REDUCE
    // On-device observer reduction: instead of storing every decimated sample, each thread keeps running
    // statistics of its observers over the samples with index >= A.sample0 (the `cutoff` of CircuitTemplate.run /
    // grid_search, frontend/template/circuit.py:550) and stores [mean, var, min, max, last][observer][instance] once at
    // the end: a 2^20-instance sweep returns 42 MB instead of 8.4 GB.  Welford's update, population variance (numpy.var).
    real r_mean[PRB_NOBS], r_m2[PRB_NOBS], r_min[PRB_NOBS], r_max[PRB_NOBS], r_last[PRB_NOBS];
    long long r_cnt = 0;
    sample = 0;
#pragma unroll
    for (int o = 0; o < PRB_NOBS; ++o) { r_mean[o] = (real)0; r_m2[o] = (real)0; r_min[o] = (real)0; r_max[o] = (real)0; r_last[o] = (real)0; }
#endif
    // The time loop runs in BURSTS: up to the next stored sample (or the end of the launch) the steps run under a 32-bit counter
    // with nothing else in the loop — the 64-bit step / decimation counters cost ~10 uniform-datapath instructions per step,
    // which are not free next to a saturated FP64 pipe (profiles/r02_issue_probe.txt).
#ifndef PRB_BURST_UNROLL
#define PRB_BURST_UNROLL
#endif
    // One thread / one warp runs (CircuitTemplate.run of a small circuit, C1) keep the plain per-step loop: the burst
    // bookkeeping is a chain of dependent uniform-datapath instructions, which a single warp cannot hide when every step is
    // stored (measured: 48 -> 132 ns per Euler step of the QIF model).  PRB_BURST_LOOP is defined for sweep kernels.
#ifdef PRB_BURST_LOOP
    long long i = 0;
    while (i < A.n_steps) {
#else
    for (long long i = 0; i < A.n_steps; ++i, ++t) {
        const int s = 0;
#endif
        if (until_store == 0) {
#ifdef PRB_REDUCE
            if (sample >= A.sample0) {
                real cur[PRB_NOBS];
                prb_record(y, cur, 0, 1);
                ++r_cnt;
                const real inv = (real)1 / (real)r_cnt;
#pragma unroll
                for (int o = 0; o < PRB_NOBS; ++o) {
                    const real d = cur[o] - r_mean[o];
                    r_mean[o] += d * inv;
                    r_m2[o] += d * (cur[o] - r_mean[o]);
                    r_min[o] = (r_cnt == 1 || cur[o] < r_min[o]) ? cur[o] : r_min[o];
                    r_max[o] = (r_cnt == 1 || cur[o] > r_max[o]) ? cur[o] : r_max[o];
                    r_last[o] = cur[o];
                }
            }
#else
#ifdef PRB_Y_SMEM
            PRB_UNROLL_NS
            for (int j = 0; j < PRB_NS; ++j) y[j] = prb_ysm_ld(T, j);
#endif
            prb_record(y, out + sample * (long long)PRB_NOBS * B, b, B);
#endif
            ++sample;
            until_store = store_step;
        }
#ifdef PRB_BURST_LOOP
        long long left = A.n_steps - i;
        if (until_store < left) left = until_store;
        if (left > 0x40000000LL) left = 0x40000000LL;
        const int burst = (int)left;                 // >= 1: until_store >= 1 here and i < n_steps
        until_store -= burst;
        PRB_BURST_UNROLL
        for (int s = 0; s < burst; ++s, ++t) {
#else
        --until_store;
        {
#endif

        PrbBad bad;
#if defined(PRB_SPEC_STEP) && !defined(PRB_Y_SMEM)
        {
            real d[PRB_NS];
            prb_step<true, false>(t, y, d, T, dt, hdt, dt6, bad);
            if (bad) {                                               // an argument left the fast range: redo the step exactly
                // local copies: arrays handed to the out-of-line function by reference would otherwise live in local
                // memory (stored every step, on the fast path too)
                real yl[PRB_NS], ynl[PRB_NS];
                PrbThread Tl = T;
#if PRB_NP > 0
                _Pragma("unroll") for (int k = 0; k < PRB_NP; ++k) Tl.p[k] = T.pbase[(long long)k * T.B + T.b];
#endif
                PRB_UNROLL_NS
                for (int j = 0; j < PRB_NS; ++j) yl[j] = y[j];
                prb_step_exact(t, yl, ynl, Tl, dt, hdt, dt6);
                PRB_UNROLL_NS
                for (int j = 0; j < PRB_NS; ++j) y[j] = ynl[j];
            } else {                                                 // both outcomes write y in place: no copies at the join
#ifdef PRB_INTEG_PAIRS
                real yo[PRB_NS];
                PRB_UNROLL_NS
                for (int j = 0; j < PRB_NS; ++j) yo[j] = y[j];
                PRB_UNROLL_NS
                for (int j = 0; j < PRB_NS; ++j)
                    y[j] = PRB_INTEG[j] >= 0 ? (yo[j] + dt * yo[PRB_INTEG[j]]) + (dt * dt6) * d[j] : yo[j] + dt6 * d[j];
#else
                PRB_UNROLL_NS
                for (int j = 0; j < PRB_NS; ++j) y[j] = y[j] + prb_step_coef(dt, hdt, dt6) * d[j];
#endif
            }
        }
#else
        real yn[PRB_NS];
#ifdef PRB_SPEC_STEP
        prb_step<true>(t, y, yn, T, dt, hdt, dt6, bad);
        if (bad) {
            real yl[PRB_NS], ynl[PRB_NS];
            PrbThread Tl = T;
#if PRB_NP > 0
            _Pragma("unroll") for (int k = 0; k < PRB_NP; ++k) Tl.p[k] = T.pbase[(long long)k * T.B + T.b];
#endif
            PRB_UNROLL_NS
            for (int j = 0; j < PRB_NS; ++j) yl[j] = y[j];
            prb_step_exact(t, yl, ynl, Tl, dt, hdt, dt6);
            PRB_UNROLL_NS
            for (int j = 0; j < PRB_NS; ++j) yn[j] = ynl[j];
        }
#else
        prb_step<false>(t, y, yn, T, dt, hdt, dt6, bad);
#endif
#ifdef PRB_Y_SMEM
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) prb_ysm_st(T, j, yn[j]);
#else
        PRB_UNROLL_NS
        for (int j = 0; j < PRB_NS; ++j) y[j] = yn[j];
#endif
#endif
#if PRB_NH > 0
        prb_hist_push(T, y, A.step0 + i + s + 1);    // DDEHistory.update((i + 1) * dt, y)
#endif
        }   // burst
#ifdef PRB_BURST_LOOP
        i += burst;
#endif
    }

#ifdef PRB_REDUCE
#pragma unroll
    for (int o = 0; o < PRB_NOBS; ++o) {
        out[(0LL * PRB_NOBS + o) * B + b] = r_mean[o];
        out[(1LL * PRB_NOBS + o) * B + b] = r_cnt > 0 ? r_m2[o] / (real)r_cnt : (real)0;
        out[(2LL * PRB_NOBS + o) * B + b] = r_min[o];
        out[(3LL * PRB_NOBS + o) * B + b] = r_max[o];
        out[(4LL * PRB_NOBS + o) * B + b] = r_last[o];
    }
#endif
#ifdef PRB_Y_SMEM
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) y[j] = prb_ysm_ld(T, j);
#endif
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) Y[(long long)j * B + b] = y[j];
    prb_thread_exit(T, A, b);
}

#endif   // fixed-step solvers

// Single rhs evaluation dy = f(t, y) for every instance (the callable handed out by get_run_func).
extern "C" __global__ void __launch_bounds__(PRB_BLOCK)
prb_rhs_eval(const __grid_constant__ prb_kernel_args A)
{
    const long long B = A.n_inst;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    PRB_STAGE_ETAB
    if (b >= B) return;
    PrbThread T;
    prb_thread_init(T, A, b);
    PRB_SET_ETAB
    const real* __restrict__ Y = (const real*)A.y;
    areal y[PRB_NS], k[PRB_NDY];
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NS; ++j) y[j] = (areal)Y[(long long)j * B + b];
#if PRB_SOLVER == PRB_SOLVER_RK45 || PRB_SOLVER == PRB_SOLVER_DOPRI5
    prb_rhs((prb_time)A.dt, y, k, T);            // float time of the evaluation travels in args.dt
#else
    prb_rhs(A.step0 + A.t0, y, k, T);
#endif
    real* __restrict__ out = (real*)A.out;
    PRB_UNROLL_NS
    for (int j = 0; j < PRB_NDY; ++j) out[(long long)j * B + b] = (real)k[j];
}
