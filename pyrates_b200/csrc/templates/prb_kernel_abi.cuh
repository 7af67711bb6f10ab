// prb_kernel_abi.cuh — device-side mirror of `prb_kernel_args` (include/prb.h).  NVRTC has no <stdint.h>,
// so the fixed-width fields are spelled with builtin types; the static_asserts pin the layout.
#ifndef PRB_KERNEL_ABI_CUH_
#define PRB_KERNEL_ABI_CUH_
struct prb_kernel_args {
    long long n_steps;
    long long store_step;
    long long step0;
    long long t0;
    long long eval0;
    long long n_inst;
    long long sample0;
    double    dt;
    void*       y;
    void*       out;
    const void* params;
    void*       slots;
    void*       rings;
    const void* tables;
    const void* consts;
    const int*  iconsts;
    void*       work;
    unsigned int* sync;
    void* const* peers;
    int         rank;
    int         world;
};
static_assert(sizeof(long long) == 8 && sizeof(int) == 4 && sizeof(void*) == 8, "LP64 expected");
static_assert(sizeof(prb_kernel_args) == 8 * 8 + 10 * 8 + 16, "prb_kernel_args layout drifted from include/prb.h");

#define PRB_SOLVER_EULER    0
#define PRB_SOLVER_HEUN_REF 1
#define PRB_SOLVER_HEUN     2
#define PRB_SOLVER_RK4      3
#define PRB_SOLVER_RK45     4
#define PRB_SOLVER_DOPRI5   5
#endif
