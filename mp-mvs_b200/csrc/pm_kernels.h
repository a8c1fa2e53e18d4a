// pm_kernels.h -- launchers of the sm_100a kernels (pm_kernels.cu), used by the C-ABI layer (pm_capi.cu).
// pm_kernels.cu is compiled twice into the library, once per arithmetic (pm_core.cuh): namespace pm_exact (the
// reference's operations bit for bit) and namespace pm_fast. PmLaunchers is the table a handle dispatches through.
#ifndef MPMVS_PM_KERNELS_H
#define MPMVS_PM_KERNELS_H
#include <cuda_runtime.h>

#include "pm_core.cuh"

#define PM_DECLARE_LAUNCHERS(NS)                                                                                                  \
    namespace NS {                                                                                                                \
    cudaError_t pm_launch_init(const PmFrame& F, const PmState& S, const PmView* gviews, unsigned long long seed, cudaStream_t st); \
    cudaError_t pm_launch_sweep(const PmFrame& F, const PmState& S, const PmView* gviews, int red, int iter, int scale,            \
                                cudaStream_t st);                                                                                  \
    cudaError_t pm_launch_finalize(const PmFrame& F, const PmState& S, cudaStream_t st);                                           \
    cudaError_t pm_launch_ncc_map(const PmFrame& F, const PmView* gviews, const pm_f4* planes, int scale, float* out,              \
                                  cudaStream_t st);                                                                                \
    cudaError_t pm_launch_ncc_bench(const PmFrame& F, const PmView* gviews, const pm_f4* planes, int scale, int taps, int nviews,  \
                                    int reps, float* out, unsigned long long* counter, cudaStream_t st);                           \
    cudaError_t pm_launch_geom_map(const PmFrame& F, const PmView* gviews, const pm_f4* planes, float* out, cudaStream_t st);      \
    }
PM_DECLARE_LAUNCHERS(pm_exact)
PM_DECLARE_LAUNCHERS(pm_fast)
#undef PM_DECLARE_LAUNCHERS

struct PmLaunchers {
    decltype(&pm_exact::pm_launch_init) init;
    decltype(&pm_exact::pm_launch_sweep) sweep;
    decltype(&pm_exact::pm_launch_finalize) finalize;
    decltype(&pm_exact::pm_launch_ncc_map) ncc_map;
    decltype(&pm_exact::pm_launch_ncc_bench) ncc_bench;
    decltype(&pm_exact::pm_launch_geom_map) geom_map;
};
// arithmetic: MPMVS_ARITH_EXACT (0) or MPMVS_ARITH_FAST (1)
inline const PmLaunchers& pm_launchers(int arithmetic) {
    static const PmLaunchers table[2] = {
        {pm_exact::pm_launch_init, pm_exact::pm_launch_sweep, pm_exact::pm_launch_finalize, pm_exact::pm_launch_ncc_map,
         pm_exact::pm_launch_ncc_bench, pm_exact::pm_launch_geom_map},
        {pm_fast::pm_launch_init, pm_fast::pm_launch_sweep, pm_fast::pm_launch_finalize, pm_fast::pm_launch_ncc_map,
         pm_fast::pm_launch_ncc_bench, pm_fast::pm_launch_geom_map}};
    return table[arithmetic ? 1 : 0];
}

// helper kernels without arithmetic of the path (defined once, in the exact build of pm_kernels.cu)
// out_fmt: 0 = float32, 1 = float16, 2 = uint8 (MPMVS_TEX_*); output rows are dense
cudaError_t pm_launch_convert(const void* in, size_t in_pitch, int in_is_u8, void* out, int out_fmt, int W, int H, cudaStream_t st);
cudaError_t pm_launch_export_depth(const pm_f4* planes, float* out, int W, int H, int pitch_floats, cudaStream_t st);
cudaError_t pm_launch_uniform_stream(unsigned long long seed, int x, int y, int n, float* out, cudaStream_t st);
// 18 tap distances + 2 reciprocals as the device evaluates them (pm_core.cuh: pm_literal_table) -> PmFrame::lit_*
cudaError_t pm_launch_literal_table(float sigma_spatial, float sigma_color, float* out20_dev, cudaStream_t st);
// MUFU.EX2 monotone over every float in [lo_negative, -0]? (the planar-prior early-out of pm_core.cuh relies on it)
cudaError_t pm_launch_ex2_monotone(float lo_negative, unsigned long long* violations_dev, cudaStream_t st);
// planar-prior stage (pm_prior.cu)
cudaError_t pm_launch_pick_vertices(const float* costs, const float* geom, int W, int H, int geom_variant, short2* out_xy,
                                    unsigned char* out_n, cudaStream_t st);
cudaError_t pm_launch_prior(const PmFrame& F, const pm_f4* planes, const int2* vxy, const int3* tris, int n_tris, pm_f4* tri_planes,
                            unsigned int* mask, pm_f4* prior, unsigned int* count, cudaStream_t st);
#endif
