// l3d_match.cuh — launch geometry + kernel declarations shared by l3d_match.cu and l3d_capi.cu
#pragma once
#include "l3d_device.cuh"
#include "../../include/l3d_capi.h"

/* geometry of k_match_topk; every knob can be overridden with -D for the tile sweep (tools/sweep_tiles.py) */
#ifndef MK_WARPS
#define MK_WARPS 8
#endif
#define MK_THREADS (32 * MK_WARPS)
#ifndef MK_RPW
#define MK_RPW 8                        /* source rows per warp */
#endif
#define MK_ROWS (MK_WARPS * MK_RPW)     /* source rows per CTA  */
#ifndef MK_TT
#define MK_TT 1024                      /* target segments per TMA stage (16 KB) */
#endif
#ifndef MK_STAGES
#define MK_STAGES 3                     /* TMA ring depth: 3072 target segments resident without reuse */
#endif
#ifndef MK_CAP
#define MK_CAP 32                       /* survivor keys kept per row before pruning to k (<= 32: one key per lane) */
#endif
#ifndef MK_MINB
#define MK_MINB 3                       /* resident CTAs per SM the register budget is tuned for */
#endif

#define DK_THREADS 256
#ifndef DK_MINB
#define DK_MINB 4                       /* resident CTAs per SM of the dense kernel (register budget 64) */
#endif
#define DK_WARPS 8
#define DK_ROWS 32                      /* MAXIMUM source rows per CTA of the dense kernel; the launch picks 8..32 so that the tile
                                           count fills whole waves of 4 CTAs per SM (l3d_dense_rows_per_cta) */
#define DK_T 4                          /* target columns per lane: a CTA covers DK_ROWS x (8*4*32 = 1024) cells */
#define DKN_ROWS 32                     /* rows per CTA of the unfiltered test kernel */

struct L3DMat3 { float m[9]; };

size_t l3d_match_smem_bytes();
size_t l3d_dense_smem_bytes();

__global__ void k_prep_segments(const float4* __restrict__ segs, const L3DViewDev* __restrict__ views, int num_views,
                                long long total, float4* __restrict__ cache);
__global__ void k_match_topk(const float4* __restrict__ segs, const float4* __restrict__ cache,
                             const L3DViewDev* __restrict__ views, const L3DPairDev* __restrict__ pairs,
                             const int2* __restrict__ tiles, int knn, float epi, int* __restrict__ counts_out,
                             l3d_match_rec* __restrict__ recs_out, const uint4* __restrict__ arcs, const L3DPairBasis* __restrict__ basis);
__global__ void k_match_topk_f64(const float4* __restrict__ segs, const float4* __restrict__ cache,
                                 const L3DViewDev* __restrict__ views, const L3DPairDev* __restrict__ pairs,
                                 const int2* __restrict__ tiles, int knn, float epi, int* __restrict__ counts_out,
                                 l3d_match_rec* __restrict__ recs_out, const double* __restrict__ cache_d, const uint4* __restrict__ arcs,
                                 const L3DPairBasis* __restrict__ basis);
__global__ void k_match_all(const float4* __restrict__ segs, const float4* __restrict__ cache,
                            const L3DViewDev* __restrict__ views, const L3DPairDev* __restrict__ pairs,
                            const int2* __restrict__ tiles, int stride, float epi, int* __restrict__ counts_out,
                            l3d_match_rec* __restrict__ recs_out, const double* __restrict__ cache_d, const uint4* __restrict__ arcs,
                            const L3DPairBasis* __restrict__ basis);
// level-1 pre-filter tables: one CTA per view pair of [first_pair, first_pair + gridDim.x): raw arcs + sort keys; after the sort
// k_arcs_gather writes the packed entries in window order
__global__ void k_pair_arcs(const float4* __restrict__ segs, const L3DViewDev* __restrict__ views, const L3DPairDev* __restrict__ pairs,
                            int first_pair, int enabled, double ext, uint4* __restrict__ raw, unsigned long long* __restrict__ keys,
                            unsigned int* __restrict__ vals, L3DPairBasis* __restrict__ basis);
__global__ void k_arcs_gather(long long n, const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ vals,
                              const uint4* __restrict__ raw, const L3DPairDev* __restrict__ pairs, int first_pair, uint4* __restrict__ out);
__global__ void k_sort_rows(int* __restrict__ counts, l3d_match_rec* __restrict__ recs, int stride, long long rows, int topk);
__global__ void k_prep_segments_f64(const float4* __restrict__ segs, const L3DViewDev* __restrict__ views, int num_views,
                                    long long total, double* __restrict__ cache);
__global__ void k_match_dense(const float4* __restrict__ ssegs, int Ns, const float4* __restrict__ tsegs, int Nt,
                              const float4* __restrict__ scache, const float4* __restrict__ tcache, L3DMat3 F, float3 Cs,
                              float3 Ct, float epi, float4* __restrict__ depths, float* __restrict__ overlaps, int rows_per_cta);
int l3d_dense_rows_per_cta(int Ns, int Nt, int num_sms);
struct L3DDenseJob {            // one view pair of a batched dense launch
    const float4* ssegs; const float4* tsegs; const float4* scache; const float4* tcache; float4* depths; float* overlaps;
    L3DMat3 F; float3 Cs, Ct; int Ns, Nt, rows_per_cta, colb; long long tile0;
};
__global__ void k_match_dense_batch(const L3DDenseJob* __restrict__ jobs, int njobs, float epi);
__global__ void k_match_dense_nofilter(const float4* __restrict__ ssegs, int Ns, const float4* __restrict__ tsegs, int Nt,
                                       const float4* __restrict__ scache, const float4* __restrict__ tcache, L3DMat3 F,
                                       float3 Cs, float3 Ct, float epi, float4* __restrict__ depths,
                                       float* __restrict__ overlaps);
__global__ void k_compact_matches(const int* __restrict__ counts, const long long* __restrict__ row_ptr,
                                  const l3d_match_rec* __restrict__ recs, int knn, long long rows,
                                  l3d_match_rec* __restrict__ out);
__global__ void k_fp32_peak(float* __restrict__ out, int iters, float a, float b);
