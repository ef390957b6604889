// Implicit-GEMM convolution kernels on tcgen05/TMEM fed by TMA (sm_100a).
//
// One "tap" of a convolution is a plain GEMM between a shifted (and, for stride 2, parity-decimated) view
// of an NHWC bf16 activation tensor and one [N x C] slab of the packed weight tensor.  A TMA box of
// (kchunk channels) x (tw x th x tb pixels) lands in shared memory as a 128-row K-major swizzled UMMA
// operand; image borders / padding come for free from TMA out-of-bounds zero fill.
//
//   fprop-type kernel  : D[pixel, n] = sum_taps sum_c  A_tap[pixel, c] * W_tap[n, c]
//       - Conv2d forward (k4 s2): 16 taps over the 4 (row parity, col parity) views of x
//       - ConvTranspose2d forward / Conv2d dgrad (k4 s2): 4 output phases x 4 taps over the plain view
//       - dense GEMM (Linear, 1x1-input ConvT): 1 tap
//   wgrad-type kernel  : dW_tap[m, n] = sum_pixels P[pixel, m] * Q_tap[pixel, n]   (both operands MN-major)
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace vg {

struct IgemmTap {
    int16_t view;   // which A tensor map (parity view) this tap reads
    int16_t dy;     // shift in view rows
    int16_t dx;     // shift in view columns
    int16_t tap_id; // wgrad: destination tap index; fprop: unused
    int32_t brow;   // fprop: first row of this tap's [N x C] slab in the packed-weight 2D map
};

struct alignas(64) IgemmParams {
    CUtensorMap amap[4];
    CUtensorMap bmap;
    IgemmTap taps[64];     // flat: taps[phase * taps_per_phase + t]
    int num_phases, taps_per_phase;
    int kchunk;    // channels per pipeline stage: 64 / 32 / 16  (row bytes 128 / 64 / 32)
    int c_chunks;  // C / kchunk
    int tw, th, tb;
    int tiles_w, tiles_h, tiles_b;
    int n_tile;    // UMMA N (multiple of 16, <= 256)
    int n_tiles;   // grid.y: N_total / n_tile
    int stages;
    int b_merged;  // the tps weight slabs of a stage are adjacent rows of the packed tensor: ONE TMA box loads them
    int tps;       // taps per pipeline stage (> 1 only when c_chunks == 1: narrow-channel layers, amortises the
                   // per-stage barrier / issue overhead over several K=16..32 slabs)
    // output: NHWC tensor, element (b, y, x, n) with y = i*osy + ay[phase], x = j*osx + ax[phase]
    void* out;
    int out_fp32;
    int out_B, out_H, out_W, out_C;
    int osy, osx;
    int ph_ay[4], ph_ax[4];
    const float* bias;  // optional [out_C]
    // split-K (single-phase launches with few output tiles and a long reduction): blockIdx.z = K slice; slices add
    // their fp32 partial tiles into `splitk_acc` (zeroed by the launcher) and a finishing kernel converts / adds bias
    int ksplit;
    float* splitk_acc;
    // Fused epilogue (bf16 output, ksplit == 1 only).  Statistics are per channel = (output column) % fuse_c and per
    // statistics group = (image index) / fuse_group_batch; a tile never straddles two groups (launcher's contract).
    //   mode 1  BatchNorm statistics of the stored (bf16-rounded) output:  sums[g][0][c] += v, sums[g][1][c] += v*v
    //   mode 2  BatchNorm backward: the accumulator is dy; out = dz = dy * act'(x*scale+shift) with x = fuse_x (the
    //           raw conv output the BatchNorm normalised), sums[g][0][c] += dz, sums[g][1][c] += dz*(x-mean)*rstd
    //   mode 3  activation backward only: out = dy * act'(x)
    //   mode 4  activation forward: out = act(acc) for ReLU / LeakyReLU layers without BatchNorm
    // CTAs accumulate in shared memory and flush once with fp32 atomics into `fuse_sums` (caller zero-initialises).
    int fuse_mode, fuse_groups, fuse_group_batch, fuse_c;
    float* fuse_sums;          // [groups][2][fuse_c]
    const void* fuse_x;        // bf16, same NHWC shape as `out`
    const float* fuse_stats;   // [groups][4][fuse_c]: mean, rstd, scale, shift
    int fuse_act;
    float fuse_slope;
    // TMA epilogue (bf16 output, no split-K, n_tile % 32 == 0): every epilogue warp moves its 32 rows x 32 columns
    // of a chunk through a 2 KB shared-memory slab - one bulk tensor store per chunk (omap[phase]: the output, or its
    // (row, column) parity view for the strided phases of an `up` contraction) and, in modes 2 / 3, one bulk tensor
    // load of the saved tensor's box (xmap[phase], same geometry) `ep_slots - 2` chunks ahead.
    CUtensorMap omap[4];
    CUtensorMap xmap[4];
    int tep;
    int ep_slots;              // slabs per epilogue warp: 2 (store double-buffer) or 3..4 (modes 2 / 3: + prefetch depth)
    int ew8;                   // 8 epilogue warps (320 threads): two per TMEM lane quadrant taking alternate chunks;
                               // for launches that own a whole SM (wide tiles), where 4 warps cannot keep up
    // Halo tiles (experiment switch VG_HALO=1; kchunk == 64, tw = 8, th = 16, tb = 1).  The taps of a phase are ordered
    // in groups of `tps` taps that read the same view at shifts within [0, hy] x [0, hx]; a pipeline stage holds ONE
    // activation tile of (th + hy) x (tw + hx) pixels and the group's `tps` weight slabs.  Group index
    // gi = (phase * taps_per_phase + first tap of the group) / tps.
    int halo;
    int halo_w, halo_h;          // tw + hx, th + hy: the TMA box of amap[] in this mode
    int halo_bytes;              // halo_w * halo_h * kchunk * 2: what one activation load delivers
    int halo_stage_bytes;        // the same rounded up to the swizzle period (1024 B)
    int16_t halo_dy[32], halo_dx[32];   // per group: origin of the halo tile relative to the output tile's (i0, j0)
    uint16_t halo_shift16[64];   // per tap (index as taps[]): (sy * halo_w + sx) * row_bytes / 16, the window's start
};

// shared memory the fused epilogue adds to a CTA
inline int igemm_fuse_smem_bytes(const IgemmParams& p) {
    if (p.fuse_mode == 1) return p.fuse_groups * 2 * p.fuse_c * 4;
    if (p.fuse_mode == 2 || p.fuse_mode == 5) return p.fuse_groups * p.fuse_c * (2 * 4 + 16);
    return 0;
}
// shared memory of the TMA epilogue's staging slabs (+ the padding that aligns them to 1024 B)
inline int igemm_epilogue_smem_bytes(const IgemmParams& p) {
    return 1024 + (p.tep ? (p.ew8 ? 8 : 4) * p.ep_slots * 2048 : 0);
}

struct alignas(64) WgradParams {
    CUtensorMap pmap;      // "P" operand (plain view), channels -> UMMA M
    CUtensorMap qmap[4];   // "Q" operand views (taps shift these), channels -> UMMA N
    IgemmTap taps[16];
    int num_taps;
    int taps_per_cta;      // accumulators resident in TMEM per CTA (taps_per_cta * n_tile <= 512)
    int merge;             // taps whose Q tiles sit side by side in one stage and are multiplied by ONE UMMA of
                           // N = merge * n_tile (<= 256): the P tile is read from smem once per `merge` taps
    int tw, th, tb;        // pixel box; tw*th*tb = kpix (multiple of 16, <= 128)
    int tiles_w, tiles_h, tiles_b;
    int p_atom_c;          // channels per MN-major atom of P: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)
    int m_atoms;           // P atoms actually loaded per CTA (UMMA M is always 128; missing atoms are ignored rows)
    int q_atom_c;          // channels per MN-major atom of Q: 64 / 32 / 16
    int n_tile;            // Q channels per CTA = UMMA N (multiple of q_atom_c and of 16, <= 256)
    int m_tiles, n_tiles;
    int splits;            // split of the pixel-tile range across CTAs
    int stages_a, stages_b;
    float* dw;             // fp32 gradient, accumulated (+=)
    long long s_m, s_n, s_tap;  // element strides of dw for (P channel, Q channel, tap)
    int m_valid, n_valid;  // channel counts actually present (rows/cols beyond are dropped)
    // Output policy.  splits == 1: every dw element is owned by exactly one CTA, which does a plain read-add-write.
    // splits > 1: each CTA stores its fp32 partial tile to `partial` ([cta][tap_local][128 rows][n_tile], coalesced
    // vector stores) and wgrad_reduce_kernel sums the splits into dw.  No atomics either way.
    float* partial;
    int debug_flags;       // bit0: skip the epilogue stores, bit1: skip TMA+MMA main loop, bit2: no MMAs, bit3: no TMA loads,
                           // bit4: epilogue stores without reading dw (profiling experiments only)
    int pair;              // launch CTA pairs (cta_group::2, M = 256): needs an even m_tiles, full 128-row M tiles, an
                           // even number of Q atoms per stage and taps_per_cta % merge == 0
    int vec4_taps;         // k*k and taps_per_cta multiples of 4, dw 16-byte aligned: dw[m][n][4 taps] moves as one float4
    int accumulate;        // 1: dw += result (default); 0: dw = result (the owner of a tile skips reading dw; split paths
                           // start from a zeroed dw / store in the reduction kernel)
    int atomic_split;      // splits > 1: every CTA adds its tile into dw with red.global.add.v4.f32 (needs vec4_taps)
};

// Host-side launchers (return cudaError_t as int).
int launch_igemm(const IgemmParams& p, cudaStream_t stream);
int igemm_total_smem(const IgemmParams& p);
int launch_wgrad(const WgradParams& p, cudaStream_t stream);
size_t wgrad_partial_bytes(const WgradParams& p);

// rank<=4 bf16 tensor map; dims/strides innermost first; strides in ELEMENTS for dims 1.. (dim 0 is contiguous).
// swizzle_bytes in {128, 64, 32}.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box, int swizzle_bytes);

}  // namespace vg
