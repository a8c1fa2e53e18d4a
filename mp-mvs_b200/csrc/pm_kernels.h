// pm_kernels.h -- launchers of the sm_100a kernels (pm_kernels.cu), used by the C-ABI layer (pm_capi.cu).
#ifndef MPMVS_PM_KERNELS_H
#define MPMVS_PM_KERNELS_H
#include <cuda_runtime.h>

#include "pm_core.cuh"

cudaError_t pm_launch_init(const PmFrame& F, const PmState& S, const PmView* gviews, unsigned long long seed, cudaStream_t st);
cudaError_t pm_launch_sweep(const PmFrame& F, const PmState& S, const PmView* gviews, int red, int iter, int scale,
                            cudaStream_t st);
cudaError_t pm_launch_finalize(const PmFrame& F, const PmState& S, cudaStream_t st);
cudaError_t pm_launch_ncc_map(const PmFrame& F, const PmView* gviews, const pm_f4* planes, int scale, float* out,
                              cudaStream_t st);
cudaError_t pm_launch_ncc_bench(const PmFrame& F, const PmView* gviews, const pm_f4* planes, int scale, int taps, int nviews, int reps,
                                float* out, unsigned long long* counter, cudaStream_t st);
cudaError_t pm_launch_geom_map(const PmFrame& F, const PmView* gviews, const pm_f4* planes, float* out, cudaStream_t st);
// out_fmt: 0 = float32, 1 = float16, 2 = uint8 (MPMVS_TEX_*); output rows are dense
cudaError_t pm_launch_convert(const void* in, size_t in_pitch, int in_is_u8, void* out, int out_fmt, int W, int H, cudaStream_t st);
cudaError_t pm_launch_export_depth(const pm_f4* planes, float* out, int W, int H, int pitch_floats, cudaStream_t st);
cudaError_t pm_launch_uniform_stream(unsigned long long seed, int x, int y, int n, float* out, cudaStream_t st);
#if PM_LITERAL_NCC == 2
// 18 tap distances + 2 reciprocals as the device evaluates them (pm_core.cuh: pm_literal_table) -> PmFrame::lit_*
cudaError_t pm_launch_literal_table(float sigma_spatial, float sigma_color, float* out20_dev, cudaStream_t st);
#endif
// planar-prior stage (pm_prior.cu)
cudaError_t pm_launch_pick_vertices(const float* costs, const float* geom, int W, int H, int geom_variant, short2* out_xy,
                                    unsigned char* out_n, cudaStream_t st);
cudaError_t pm_launch_prior(const PmFrame& F, const pm_f4* planes, const int2* vxy, const int3* tris, int n_tris, pm_f4* tri_planes,
                            unsigned int* mask, pm_f4* prior, unsigned int* count, cudaStream_t st);
#endif
