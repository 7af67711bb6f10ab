/* prb.h — C ABI of libprb (pyrates_b200/csrc/prb.cpp): JIT + launch of PyRates vector-field kernels
 * on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the numerical-integration hot path of PyRates.  The entry points
 * replace, for `backend='cuda'`, what the reference does in Python:
 *
 *   prb_compile / prb_module_*   <->  BaseBackend.generate_func   pyrates/backend/base/base_backend.py:537-578
 *                                     (compile the generated vector field; here: CUDA C -> cubin via NVRTC)
 *   prb_run                      <->  BaseBackend.run / _solve / _solve_euler / _solve_heun
 *                                     pyrates/backend/base/base_backend.py:596-611, 693-707, 712-769
 *                                     (fixed-step time loop + state recording, fused into the kernel)
 *   prb_malloc / prb_memcpy_*    <->  BaseBackend.get_var (argument arrays handed to the generated function)
 *                                     pyrates/backend/base/base_backend.py:326-340
 *   prb_last_error               <->  PyRatesException   pyrates/backend/__init__.py:42
 *
 * Plain C types only (pointers + sizes); no torch / numpy types cross this boundary.  Every function
 * returns 0 on success or a negative prb_status; the message is available from prb_last_error().
 * The library has NO CPU execution path: without a CUDA driver/GPU every compute entry point fails
 * with PRB_ERR_NO_DEVICE.  prb_compile alone works without a GPU (NVRTC cross-compiles to sm_100a).
 */
#ifndef PRB_H_
#define PRB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRB_ABI_VERSION 2

typedef enum prb_status {
    PRB_OK = 0,
    PRB_ERR_INVALID = -1,      /* bad argument */
    PRB_ERR_COMPILE = -2,      /* NVRTC failed; see prb_module_log / prb_last_error */
    PRB_ERR_NO_DEVICE = -3,    /* libcuda / GPU not available */
    PRB_ERR_CUDA = -4,         /* driver API error */
    PRB_ERR_NOT_FOUND = -5,    /* kernel name not in module */
    PRB_ERR_LAUNCH = -6        /* launch configuration rejected (e.g. cooperative grid too large) */
} prb_status;

/* Solver ids.  HEUN_REF is the reference's `solver='heun'`, whose in-place rhs buffer aliasing makes
 * the update y <- y + dt*f(y + dt*f(y))  (base_backend.py:763-765; SURVEY.md headline 4).
 * HEUN is the textbook explicit trapezoid; RK4 the classic 4-stage scheme (no reference counterpart). */
enum { PRB_SOLVER_EULER = 0, PRB_SOLVER_HEUN_REF = 1, PRB_SOLVER_HEUN = 2, PRB_SOLVER_RK4 = 3,
       /* the reference's solver='scipy': solve_ivp(method='RK45', first_step=dt, t_eval=times), base_backend.py:802-812;
        * options (rtol, atol, t_start, t_bound, first_step, sample spacing, max_step, max attempts) are 8 doubles behind
        * args.consts, args.n_steps = samples, per-instance status words behind args.sync */
       PRB_SOLVER_RK45 = 4,
       /* the reference's solver='scipy' on a delay-differential model: scipy.integrate.ode('dopri5', first_step=dt,
        * nsteps=50000) restarted per sample interval, accepted steps appended to the state history
        * (BaseBackend._solve_scipy_dde, base_backend.py:771-800); options behind args.consts: [0] rtol [1] atol [2] t_start
        * [4] first_step [5] sample spacing [7] nsteps [8] history capacity [9..] float64 initial condition of the history;
        * args.work = history ring, status per instance behind args.sync (3 = ring too small, caller retries larger) */
       PRB_SOLVER_DOPRI5 = 5 };

/* Kernel argument block.  Passed BY VALUE as the single __grid_constant__ parameter of every generated
 * kernel; the generated CUDA source embeds the identical definition (pyrates_b200/csrc/prb_kernel_abi.h).
 * All pointers are DEVICE pointers owned by the caller.  Layouts (B = n_inst, "real" = float|double):
 *   y      [n_state][B]            state, structure-of-arrays; in: initial state, out: final state
 *   out    [n_samples][n_obs][B]   recorded observers; row k holds the state at step k*store_step,
 *                                  sampled BEFORE the update (base_backend.py:730-733)
 *   params [n_params][B]           per-instance (swept) parameters
 *   slots  [n_slots][B]            persistent mutable scalars (non-ring buffers mutated by the rhs); adaptive NETWORK kernels
 *                                  (PRB_SOLVER_RK45, cooperative) read their 8 option doubles here and use [8..12] as
 *                                  error-norm accumulators / status / attempt count
 *   rings  [ring elems][B]         delay ring buffers, physical storage, circular addressing
 *   tables [..]                    timed-input tables u[t] shared by all instances (circuit.py:1711-1719)
 *   consts / iconsts               packed real / int32 constant arrays of network-mode kernels (W, idx; for the tcgen05
 *                                  sweep kernels also W split into hi / lo K-major 128 x 32 tiles); adaptive INSTANCE
 *                                  kernels read their 8 option doubles behind `consts`
 *   work                           network-mode scratch (stage vectors, materialised intermediates); instance kernels of
 *                                  delay-differential models keep their state history here: fixed-step solvers a ring
 *                                  [depth][n_hist][B] of past states, PRB_SOLVER_DOPRI5 times [cap][B] + states [cap][n_hist][B]
 *   out (reduce kernels)           [5][n_obs][B]: mean, var, min, max, last of the observers over the samples >= sample0
 *   sync                           network-mode barrier words, 64-bit each (zero-initialised by the caller):
 *                                  [0] local CTA arrivals, [2] error flag (a bounded wait timed out; peers set it remotely)
 *   peers                          multi-GPU network mode (rows sharded, one process per GPU): device array of 2*world
 *                                  pointers {mirror[0..world), sync[0..world)} of all ranks (IPC-mapped; entry `rank` is the
 *                                  caller's own).  A mirror holds two parities of one 16-byte packet slot per work / ring
 *                                  element + 32 heartbeat slots: owners send {lo32, tag, hi32, tag} packets into their peers'
 *                                  mirrors, receivers unpack them into their replicas before each grid barrier
 *                                  (csrc/templates/net_kernel.cuh: prb_ll_*); zero-initialised by the caller before a launch
 */
typedef struct prb_kernel_args {
    int64_t n_steps;      /* solver steps to perform in this launch */
    int64_t store_step;   /* record every store_step-th step */
    int64_t step0;        /* index of this launch's first step (store cadence + `t` continue across launches) */
    int64_t t0;           /* value of the generated function's `t` at step 0: t = step + t0 (base_backend.py:734) */
    int64_t eval0;        /* rhs evaluations performed before this launch (ring-buffer head position) */
    int64_t n_inst;       /* B */
    int64_t sample0;      /* first output row written by this launch */
    double  dt;
    void*       y;
    void*       out;
    const void* params;
    void*       slots;
    void*       rings;
    const void* tables;
    const void* consts;
    const int32_t* iconsts;
    void*       work;
    unsigned int* sync;
    void* const* peers;   /* NULL unless world > 1 */
    int32_t     rank;
    int32_t     world;
} prb_kernel_args;

/* Launch descriptor for prb_run. */
typedef struct prb_run_desc {
    uint32_t struct_size;   /* = sizeof(prb_run_desc) (ABI check) */
    int32_t  cooperative;   /* 0: plain launch (instance kernels); 1: cooperative persistent grid (network kernels) */
    uint32_t grid_dim;      /* 0 = library chooses: ceil(n_inst/block) or SMs x occupancy for cooperative */
    uint32_t block_dim;
    uint32_t smem_bytes;    /* dynamic shared memory (tcgen05 operand ring: up to 192 KB; DMMA / FMA GEMM tiles: up to 75 KB) */
    uint32_t reserved;
    void*    stream;        /* CUstream or NULL */
    prb_kernel_args args;
} prb_run_desc;

typedef struct prb_device_info {
    char     name[128];
    int32_t  cc_major, cc_minor;
    int32_t  sm_count;
    int32_t  max_smem_per_block_optin;
    int32_t  l2_bytes;
    int32_t  clock_khz;
    int64_t  total_mem;
} prb_device_info;

typedef struct prb_kernel_info {
    int32_t regs;
    int32_t static_smem;
    int32_t local_bytes;          /* per-thread local memory (spills) */
    int32_t max_threads;
    int32_t max_blocks_per_sm;    /* for the block_dim/smem passed to prb_kernel_info */
} prb_kernel_info;

typedef struct prb_module prb_module;

/* --- library / errors --------------------------------------------------------------------------- */
int         prb_abi_version(void);
const char* prb_last_error(void);             /* thread-local message of the last failing call */
int64_t     prb_launch_count(void);           /* kernels launched by this library since load */
int         prb_nvrtc_version(int* major, int* minor);   /* toolkit that prb_compile uses (part of the callers' cubin cache key) */

/* --- device ------------------------------------------------------------------------------------- */
int prb_device_count(int* count);
int prb_set_device(int ordinal);              /* retains the primary context; must precede compute calls */
int prb_get_device_info(prb_device_info* info);
int prb_device_synchronize(void);

/* --- JIT ---------------------------------------------------------------------------------------- */
/* NVRTC-compile CUDA C `src` for `arch` (e.g. "sm_100a") with extra `opts`.  Works without a GPU.
 * On failure *out is still set when a log is available (free it with prb_module_free). */
int prb_compile(const char* src, const char* name, const char* arch,
                const char* const* opts, int n_opts, prb_module** out);
int prb_module_from_cubin(const void* cubin, size_t size, prb_module** out);
int prb_module_cubin(const prb_module* m, const void** data, size_t* size);
int prb_module_log(const prb_module* m, const char** log);
int prb_module_free(prb_module* m);
int prb_kernel_get_info(prb_module* m, const char* kernel, uint32_t block_dim, uint32_t smem_bytes,
                        prb_kernel_info* info);

/* --- run ---------------------------------------------------------------------------------------- */
/* Launch generated kernel `kernel` of module `m`: performs args.n_steps fixed-step solver steps with the
 * rhs fused in, recording observers.  Asynchronous on desc->stream. */
int prb_run(prb_module* m, const char* kernel, const prb_run_desc* desc);

/* --- memory / streams / events (thin driver-API wrappers so callers need no CUDA binding) --------- */
int prb_malloc(void** dptr, size_t bytes);
int prb_free(void* dptr);
int prb_memset(void* dptr, int byte, size_t bytes, void* stream);
int prb_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int prb_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
int prb_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
/* CUDA IPC: share a prb_malloc'ed buffer with the other ranks of a multi-GPU network run (one process per GPU);
 * the opening process gets a peer pointer that kernels may load/store over NVLink. */
#define PRB_IPC_HANDLE_BYTES 64
int prb_ipc_get_handle(void* dptr, unsigned char* handle /* [PRB_IPC_HANDLE_BYTES] */);
int prb_ipc_open_handle(const unsigned char* handle, void** dptr);
int prb_ipc_close_handle(void* dptr);
int prb_host_alloc(void** hptr, size_t bytes);       /* pinned host memory */
int prb_host_free(void* hptr);
/* page-lock caller-allocated host memory (e.g. an anonymous mapping backed by transparent huge pages: an 8 GB result
 * buffer registers several times faster than cuMemHostAlloc builds it from 4 KB pages) */
int prb_host_register(void* hptr, size_t bytes);
int prb_host_unregister(void* hptr);
int prb_stream_create(void** stream);
int prb_stream_destroy(void* stream);
int prb_stream_synchronize(void* stream);
int prb_stream_wait_event(void* stream, void* event);   /* later work on `stream` waits for `event` */
int prb_event_create(void** event);
int prb_event_destroy(void* event);
int prb_event_record(void* event, void* stream);
int prb_event_synchronize(void* event);
int prb_event_elapsed_ms(void* start, void* stop, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* PRB_H_ */
