// tcgen05 / TMEM / TMA implicit-GEMM kernels (sm_100a only).  See igemm_umma.cuh for the formulation.
//
// Warp roles in a 192-thread CTA:
//   warp 0 (one lane)  : TMA producer  - fills the smem ring, arms `full[s]` with expect_tx
//   warp 1             : TMEM allocator; one lane issues tcgen05.mma and commits `empty[s]` / `tmem_full`
//   warps 2..5         : epilogue - tcgen05.ld the fp32 accumulator (lane quadrant = warp % 4), convert, store
#include "igemm_umma.cuh"
#include "ptx.cuh"
#include "pdl.cuh"

#include <algorithm>
#include <cstdio>
#include <mutex>

namespace vg {

// tcgen05.fence::after_thread_sync after every "stage full" wait of the MMA-issuing thread: the operands arrive through
// the async proxy (TMA -> mbarrier complete_tx) and are consumed by tcgen05.mma in the same proxy, so the fence orders
// nothing that the mbarrier does not already order.  -DVG_STAGE_FENCE=1 restores it (experiment switch).
#ifndef VG_STAGE_FENCE
#define VG_STAGE_FENCE 0
#endif
static constexpr int kIgemmThreads = 192;        // TMA warp, MMA warp, 4 epilogue warps
static constexpr int kIgemmMaxThreads = 320;     // ... or 8 epilogue warps (IgemmParams::ew8)
static constexpr int kMaxStages = 24;
static constexpr int kSlabBytes = 2048;          // one epilogue staging slab: 32 rows x 32 bf16 columns
static constexpr int kBarrierBytes = (4 * kMaxStages + 4 + 32) * 8 + 16;

// Column totals over the 32 lanes of a warp: v[j] is this lane's (= this output row's) value of column j.  A
// transposing butterfly (16+8+4+2+1 = 31 shuffles instead of 32 x 5) leaves the total of column `lane` in the
// returned value.  Destroys v.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int j = 0; j < o; ++j) {
            const float send = upper ? v[j] : v[j + o];
            const float keep = upper ? v[j + o] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

__device__ __forceinline__ float fuse_act_grad(float z, int act, float slope) {
    // VgAct: 1 = ReLU, 2 = LeakyReLU; anything else is the identity (the launcher admits only these)
    return (act == 1) ? (z > 0.f ? 1.f : 0.f) : ((act == 2) ? (z > 0.f ? 1.f : slope) : 1.f);
}

__device__ __forceinline__ uint32_t tmem_cols_for(int n) {
    uint32_t c = 32;
    while (c < static_cast<uint32_t>(n)) c <<= 1;
    return c;
}

// ------------------------------------------------------------------------------------------------
// fprop-type kernel: A = activation views (K-major), B = packed weights (K-major), D -> NHWC output
// ------------------------------------------------------------------------------------------------
template <int FMODE, bool HALO = false>
__global__ void __launch_bounds__(kIgemmMaxThreads) igemm_fprop_kernel(const __grid_constant__ IgemmParams p) {
    // FMODE = IgemmParams::fuse_mode (compile-time: the plain kernel carries none of the fused-epilogue code).
    // HALO  = IgemmParams::halo: the `tps` taps of a stage are the shifted windows of ONE (th+hy) x (tw+hx) halo tile
    //         (tw = 8, tb = 1: window row m = pixel (m / 8, m % 8) = halo row shift + (m / 8) * halo_w + m % 8, i.e. an
    //         8-row-group stride of halo_w rows - tests/native/umma_halo.cu shows the tensor core reads exactly those
    //         rows when the descriptor's base-offset field is 0).  One activation load per stage instead of `tps`.
    // Persistent: each CTA walks work items (M tile, N tile, phase, K slice) with stride gridDim.x.  The smem ring
    // runs continuously across items and the accumulator is double-buffered in TMEM, so the epilogue of item i
    // overlaps the loads and MMAs of item i+1 and the per-CTA prologue (TMEM alloc, barrier init) is paid once.
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    // (warp index through a shuffle: the compiler then KNOWS it is warp-uniform and keeps the role branches, and every
    // address / descriptor computed inside them, on the uniform datapath instead of "waterfall" R2UR loops)
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int row_bytes = p.kchunk * 2;
    const int tps = p.tps > 1 ? p.tps : 1;
    const int a_sub = 128 * row_bytes, b_sub = p.n_tile * row_bytes;   // one tap's operand tiles
    const int a_stage = HALO ? p.halo_stage_bytes : tps * a_sub;
    const int b_stage = tps * b_sub;
    const int stages = p.stages;
    uint8_t* sA = smem;
    uint8_t* sB = smem + stages * a_stage;
    // epilogue staging slabs (TMA epilogue): [epilogue warp][slot][32 rows x 64 B], SWIZZLE_64B, 1024-byte aligned
    uint8_t* sE = sB + stages * b_stage;
    sE += (1024u - (smem_u32(sE) & 1023u)) & 1023u;
    const int n_ewarps = p.ew8 ? 8 : 4;
    uint64_t* full = reinterpret_cast<uint64_t*>(sE + (p.tep ? n_ewarps * p.ep_slots * kSlabBytes : 0));
    uint64_t* empty = full + kMaxStages;
    uint64_t* tmem_full = empty + kMaxStages;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* ebar = tmem_empty + 2;               // [8 epilogue warps][4 slots]: "saved-tensor box has landed"
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ebar + 32);
    // fused-epilogue scratch: per-CTA channel sums [groups][2][C], then (mode 2) per-channel (mean, rstd, scale, shift)
    constexpr int fmode = FMODE;
    const int fC = p.fuse_c;
    float* s_sums = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + kBarrierBytes);
    float4* s_prm = reinterpret_cast<float4*>(s_sums + p.fuse_groups * 2 * fC);
    if (fmode == 1 || fmode == 2)
        for (int i = threadIdx.x; i < p.fuse_groups * 2 * fC; i += blockDim.x) s_sums[i] = 0.f;

    const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    const int total_items = m_tiles * p.n_tiles * p.num_phases * ksplit;
    const int total_iters = p.taps_per_phase / tps * p.c_chunks;
    const uint32_t ncols = tmem_cols_for(2 * p.n_tile);

    if (warp == 0 && lane == 0) {
        for (int v = 0; v < 4; ++v) tma_prefetch_desc(&p.amap[v]);
        tma_prefetch_desc(&p.bmap);
        if (p.tep)
            for (int v = 0; v < p.num_phases; ++v) {
                tma_prefetch_desc(&p.omap[v]);
                if (fmode == 2 || fmode == 3) tma_prefetch_desc(&p.xmap[v]);
            }
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], n_ewarps);      // one arrival per epilogue warp
        }
        for (int i = 0; i < 32; ++i) mbar_init(&ebar[i], 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, ncols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    // everything above (barriers, tensor memory, descriptor prefetch) overlapped the previous kernel's tail; from here
    // on this grid reads what that kernel wrote (pdl.cuh)
    pdl_enter();
    if (fmode == 2 || fmode == 5) {
        for (int i = threadIdx.x; i < p.fuse_groups * fC; i += blockDim.x) {
            const int g = i / fC, c = i - g * fC;
            const float* st = p.fuse_stats + static_cast<size_t>(g) * 4 * fC;
            s_prm[i] = make_float4(__ldg(st + c), __ldg(st + fC + c), __ldg(st + 2 * fC + c), __ldg(st + 3 * fC + c));
        }
        __syncthreads();
    }

    // item -> coordinates (m fastest: consecutive CTAs share the same weight tile)
    auto decode = [&](int item, int& i0, int& j0, int& b0, int& n0, int& phase, int& it_begin, int& iters) {
        int t = item % m_tiles;
        int r = item / m_tiles;
        const int n_idx = r % p.n_tiles;
        const int z = r / p.n_tiles;
        j0 = (t % p.tiles_w) * p.tw;
        t /= p.tiles_w;
        i0 = (t % p.tiles_h) * p.th;
        b0 = (t / p.tiles_h) * p.tb;
        n0 = n_idx * p.n_tile;
        phase = z / ksplit;
        const int kslice = z % ksplit;
        it_begin = static_cast<int>(static_cast<long long>(total_iters) * kslice / ksplit);
        iters = static_cast<int>(static_cast<long long>(total_iters) * (kslice + 1) / ksplit) - it_begin;
    };

    // Producer and issuer warps run their loops with ALL 32 lanes (warp-uniform control flow and address arithmetic)
    // and only elect lane 0 around the barrier-arrive / TMA / tcgen05 instructions themselves.  With the whole loop
    // inside `if (lane == 0)` every operand lives in per-thread registers and each tcgen05.mma costs the thread ~77
    // cycles, each 4-D TMA ~300 (trace build, round 1): the operands have to be moved to uniform registers one by one.
    const bool leader = lane == 0;
    if (warp == 0) {
        {
            int s = 0;
            uint32_t par = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                int i0, j0, b0, n0, phase, it_begin, iters;
                decode(item, i0, j0, b0, n0, phase, it_begin, iters);
                const IgemmTap* taps = &p.taps[phase * p.taps_per_phase];
                int tap_i = it_begin / p.c_chunks * tps, c = it_begin % p.c_chunks;
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(&empty[s], par ^ 1);
                    // (elect.sync, not `lane == 0`: ptxas then knows exactly one thread runs the block and emits the
                    // TMA / tcgen05 instructions straight-line instead of one ELECT + BRA.U.ANY loop per instruction)
                    if (elect_one()) {
                        mbar_expect_tx(&full[s], (HALO ? p.halo_bytes : a_stage) + b_stage);
                        if (HALO) {
                            // the group's halo tile: origin = the smallest shift of its taps (IgemmParams::halo_dy / dx)
                            const int gi = (phase * p.taps_per_phase + tap_i) / tps;
                            tma_load_4d(sA + s * a_stage, &p.amap[taps[tap_i].view], &full[s], c * p.kchunk,
                                        j0 + p.halo_dx[gi], i0 + p.halo_dy[gi], b0);
                        }
                        for (int t = 0; t < tps; ++t) {
                            const IgemmTap tap = taps[tap_i + t];
                            if (!HALO)
                                tma_load_4d(sA + s * a_stage + t * a_sub, &p.amap[tap.view], &full[s], c * p.kchunk,
                                            j0 + tap.dx, i0 + tap.dy, b0);
                            if (!p.b_merged)
                                tma_load_2d(sB + s * b_stage + t * b_sub, &p.bmap, &full[s], c * p.kchunk, tap.brow + n0);
                        }
                        if (p.b_merged)
                            tma_load_2d(sB + s * b_stage, &p.bmap, &full[s], c * p.kchunk, taps[tap_i].brow + n0);
                    }
                    __syncwarp();
                    if (++c == p.c_chunks) { c = 0; tap_i += tps; }
                    if (++s == stages) { s = 0; par ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        {
            const uint32_t idesc = make_idesc_bf16(128, p.n_tile, 0, 0);
            const uint32_t layout = p.kchunk == 64 ? 2u : (p.kchunk == 32 ? 4u : 6u);
            const uint32_t sbo = 8 * row_bytes;
            const uint32_t sbo_a = sbo;     // (halo windows start on swizzle-atom boundaries: standard descriptors)
            const int ksteps = p.kchunk / 16;
            // descriptors of stage 0 / k-step 0; per instruction only the start-address field (lo word) moves
            const uint64_t a_desc0 = make_smem_desc(smem_u32(sA), 0, sbo_a, layout);
            const uint64_t b_desc0 = make_smem_desc(smem_u32(sB), 0, sbo, layout);
            const uint32_t a_hi = static_cast<uint32_t>(a_desc0 >> 32), b_hi = static_cast<uint32_t>(b_desc0 >> 32);
            const uint32_t a_lo0 = static_cast<uint32_t>(a_desc0), b_lo0 = static_cast<uint32_t>(b_desc0);
            const uint32_t a_step = a_stage >> 4, b_step = b_stage >> 4;
            const uint32_t a_t = a_sub >> 4, b_t = b_sub >> 4;
            int s = 0;
            uint32_t par = 0, a_lo = a_lo0, b_lo = b_lo0;
            int li = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
                int i0, j0, b0, n0, phase, it_begin, iters;
                decode(item, i0, j0, b0, n0, phase, it_begin, iters);
                const int acc = li & 1;
                mbar_wait(&tmem_empty[acc], ((li >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * p.n_tile;
                // halo: the stage's taps differ in where their window starts inside the tile (halo_shift16, in
                // 16-byte units), which follows the tap group -> track it like the producer does
                int tap_i = HALO ? it_begin / p.c_chunks * tps : 0, hc = HALO ? it_begin % p.c_chunks : 0;
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(&full[s], par);
#if VG_STAGE_FENCE
                    tc_fence_after();
#endif
                    const bool issuer = elect_one();     // (elect.sync: ptxas then KNOWS one thread runs the block)
                    if (HALO) {
                        if (issuer) {
                            const uint16_t* sh = &p.halo_shift16[phase * p.taps_per_phase + tap_i];
                            for (int t = 0; t < tps; ++t)
                                for (int k = 0; k < ksteps; ++k)
                                    umma_bf16_lohi(d_tmem, a_lo + sh[t] + 2 * k, a_hi, b_lo + t * b_t + 2 * k, b_hi,
                                                   idesc, (it | t | k) != 0);
                            umma_commit(&empty[s]);
                        }
                        if (++hc == p.c_chunks) { hc = 0; tap_i += tps; }
                    } else if (issuer) {
                        if (ksteps == 4) {
                            umma_bf16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, it != 0);
                            umma_bf16_lohi(d_tmem, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1);
                            umma_bf16_lohi(d_tmem, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1);
                            umma_bf16_lohi(d_tmem, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1);
                        } else {
                            for (int t = 0; t < tps; ++t)
                                for (int k = 0; k < ksteps; ++k)
                                    umma_bf16_lohi(d_tmem, a_lo + t * a_t + 2 * k, a_hi, b_lo + t * b_t + 2 * k, b_hi,
                                                   idesc, (it | t | k) != 0);
                        }
                        umma_commit(&empty[s]);
                    }
                    __syncwarp();
                    a_lo += a_step;
                    b_lo += b_step;
                    if (++s == stages) { s = 0; par ^= 1; a_lo = a_lo0; b_lo = b_lo0; }
                }
                if (elect_one()) umma_commit(&tmem_full[acc]);   // (with zero iterations this arrives immediately)
                __syncwarp();
            }
        }
    } else {
        // Epilogue warps.  A warp may read only the TMEM lane quadrant (warp % 4); with p.ew8 (one CTA per SM, wide
        // tiles) two warps share a quadrant and take alternate 32-column chunks.
        const int ew = warp - 2;
        const int q = warp & 3;
        const int c_begin = (ew >> 2) * 32, c_step = p.ew8 ? 64 : 32;
        const int row = q * 32 + lane;
        const int w_l = row % p.tw;
        const int h_l = (row / p.tw) % p.th;
        const int b_l = row / (p.tw * p.th);
        const __nv_bfloat16* fx = static_cast<const __nv_bfloat16*>(p.fuse_x);
        // derivative of the fused layer's activation on the negative side (ReLU 0, LeakyReLU slope, identity 1)
        const float neg_slope = p.fuse_act == 1 ? 0.f : (p.fuse_act == 2 ? p.fuse_slope : 1.f);
        constexpr bool mode23 = (FMODE == 2 || FMODE == 3);
        // TMA epilogue (p.tep): the warp's 32 rows x 32 columns of a chunk go through a 2 KB shared-memory slab
        // (64-byte rows, SWIZZLE_64B: conflict-free 16-byte accesses) and leave with ONE bulk tensor store; in modes
        // 2 / 3 the same slab first receives the saved tensor's box by TMA, `ep_slots - 2` chunks ahead, and the
        // result overwrites it in place.  Per-lane 16-byte global accesses at pixel stride cost 32 L1 wavefronts per
        // instruction and made the LSU pipe the busiest unit of the narrow-tile launches (ncu, round 2).
        const bool tep = p.tep != 0;
        const int S = p.ep_slots;
        const bool xtma = tep && mode23 && S >= 3;      // else (wide tiles, S = 1) the saved tensor comes per lane
        const uint32_t slab0 = smem_u32(sE) + static_cast<uint32_t>(ew * S) * kSlabBytes;
        const uint32_t lrow = slab0 + static_cast<uint32_t>(lane) * 64u;       // this lane's row inside slot 0
        const uint32_t sw = (lane >> 1) & 3;                                   // SWIZZLE_64B: unit ^= (row >> 1) & 3
        uint64_t* xbar = ebar + ew * 4;
        const int r0 = q * 32;
        const int sub_w = r0 % p.tw, sub_h = (r0 / p.tw) % p.th, sub_b = r0 / (p.tw * p.th);
        int slot = 0;
        uint32_t xpar = 0;
        // prefetch cursor of the saved-tensor boxes (warp-uniform: every lane tracks it, one elected lane issues)
        int pf_item = blockIdx.x, pf_c = c_begin, pf_slot = 0, pf_j = 0, pf_i = 0, pf_b = 0, pf_n0 = 0, pf_phase = 0;
        auto pf_setup = [&]() {
            if (pf_item >= total_items) return;
            int i0, j0, b0, n0, phase, it_begin, iters;
            decode(pf_item, i0, j0, b0, n0, phase, it_begin, iters);
            pf_j = j0 + sub_w; pf_i = i0 + sub_h; pf_b = b0 + sub_b; pf_n0 = n0; pf_phase = phase;
        };
        auto pf_issue = [&]() {
            if (pf_item >= total_items) return;
            if (elect_one()) {
                mbar_expect_tx(&xbar[pf_slot], kSlabBytes);
                tma_load_4d(sE + (ew * S + pf_slot) * kSlabBytes, &p.xmap[pf_phase], &xbar[pf_slot], pf_n0 + pf_c, pf_j,
                            pf_i, pf_b);
            }
            if (++pf_slot == S) pf_slot = 0;
            pf_c += c_step;
            if (pf_c >= p.n_tile) { pf_c = c_begin; pf_item += gridDim.x; pf_setup(); }
        };
        if (xtma) {
            pf_setup();
            for (int i = 0; i < S - 2; ++i) pf_issue();
        }
        __syncwarp();
        int li = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
            int i0, j0, b0, n0, phase, it_begin, iters;
            decode(item, i0, j0, b0, n0, phase, it_begin, iters);
            const int acc = li & 1;
            const int b = b0 + b_l;
            const int y = (i0 + h_l) * p.osy + p.ph_ay[phase];
            const int x = (j0 + w_l) * p.osx + p.ph_ax[phase];
            const bool valid = (b < p.out_B) && (y < p.out_H) && (x < p.out_W);
            const size_t off = ((static_cast<size_t>(b) * p.out_H + y) * p.out_W + x) * p.out_C + n0;
            const bool all_valid = __all_sync(0xffffffffu, valid);
            int grp = 0;
            if (fmode == 1 || fmode == 2) grp = min(b0 / p.fuse_group_batch, p.fuse_groups - 1);   // tile-uniform
            float* gs0 = s_sums + grp * 2 * fC;
            float* gs1 = gs0 + fC;

            // mode 2/3 without the TMA epilogue: this row's slice of the saved conv output, one chunk ahead
            uint4 xq[4] = {};
            auto load_x = [&](int c) {
                if (valid) {
                    const uint4* src = reinterpret_cast<const uint4*>(fx + off + c);
                    xq[0] = __ldg(src);
                    xq[1] = __ldg(src + 1);
                    if (c + 16 < p.n_tile) {
                        xq[2] = __ldg(src + 2);
                        xq[3] = __ldg(src + 3);
                    }
                }
            };
            if (mode23 && !xtma) load_x(c_begin);

            mbar_wait(&tmem_full[acc], (li >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.n_tile;
            for (int c = c_begin; c < p.n_tile; c += c_step) {
                uint32_t v[32];
                const bool full_chunk = (p.n_tile - c) >= 32;       // else a 16-column tail
                if (full_chunk) {
                    tmem_ld_32x32(taddr + c, v);
                } else {
                    uint32_t h[16];
                    tmem_ld_32x16(taddr + c, h);
#pragma unroll
                    for (int j = 0; j < 16; ++j) { v[j] = h[j]; v[j + 16] = 0; }
                }
                const uint32_t my = lrow + static_cast<uint32_t>(slot) * kSlabBytes;
                if (tep) {
                    // the slab this chunk uses was last read by the store issued two chunks ago (S = 2) / is about to
                    // be re-filled for the chunk S - 2 ahead: at most the newest store may still be reading
                    // (bulk async-groups belong to the thread that committed them: lane 0 stores, lane 0 waits)
                    if (lane == 0) {
                        if (S == 1) bulk_wait_read<0>(); else bulk_wait_read<1>();
                    }
                    __syncwarp();
                    if (xtma) pf_issue();
                    if (xtma) {
                        mbar_wait(&xbar[slot], xpar);
#pragma unroll
                        for (int k = 0; k < 4; ++k) xq[k] = ld_shared_v4(my + ((static_cast<uint32_t>(k) ^ sw) << 4));
                    }
                }
                tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (fmode == 0 && ksplit > 1) {
                    if (valid && iters > 0) {
                        float* dst = p.splitk_acc + off + c;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < 16 || full_chunk) atomicAdd(dst + j, f[j]);
                    }
                    continue;
                }
                if (p.bias != nullptr) {
                    const float* bp = p.bias + n0 + c;
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] += __ldg(bp + j);
                    if (full_chunk) {
#pragma unroll
                        for (int j = 16; j < 32; ++j) f[j] += __ldg(bp + j);
                    }
                }
                if (fmode == 0 && p.out_fp32) {
                    if (valid) {
                        float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + off + c);
#pragma unroll
                        for (int j = 0; j < 4; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                        if (full_chunk) {
#pragma unroll
                            for (int j = 4; j < 8; ++j)
                                dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                        }
                    }
                    continue;
                }
                // fused statistics: fuse_c and n_tile are multiples of 32, so a chunk never wraps around the channels
                int ch0 = 0;
                if (fmode == 1 || fmode == 2 || fmode == 5) ch0 = (n0 + c) % fC;
                float xv[32];
                if (mode23) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t w4[4] = {xq[j].x, xq[j].y, xq[j].z, xq[j].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            xv[8 * j + 2 * i] = __uint_as_float(w4[i] << 16);
                            xv[8 * j + 2 * i + 1] = __uint_as_float(w4[i] & 0xFFFF0000u);
                        }
                    }
                    if (!xtma && c + c_step < p.n_tile) load_x(c + c_step);
                    if (fmode == 2) {
                        const float4* pr_base = s_prm + grp * fC + ch0;
#pragma unroll
                        for (int jb = 0; jb < 32; jb += 8) {
                            float4 pr[8];                       // (mean, rstd, scale, shift) of 8 columns, loaded together
#pragma unroll
                            for (int j = 0; j < 8; ++j) pr[j] = pr_base[jb + j];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float z = fmaf(xv[jb + j], pr[j].z, pr[j].w);
                                f[jb + j] *= z > 0.f ? 1.f : neg_slope;
                                xv[jb + j] = (xv[jb + j] - pr[j].x) * pr[j].y;          // xhat
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] *= xv[j] > 0.f ? 1.f : neg_slope;
                    }
                }
                if (fmode == 5) {
                    // eval-mode BatchNorm folded into the forward epilogue: out = act(conv * scale[c] + shift[c]) from
                    // the fp32 accumulator (the stand-alone pass it replaces re-read a bf16-rounded conv output)
                    const float4* pr_base = s_prm + ch0;
#pragma unroll
                    for (int jb = 0; jb < 32; jb += 8) {
                        float4 pr[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) pr[j] = pr_base[jb + j];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float z = fmaf(f[jb + j], pr[j].z, pr[j].w);
                            f[jb + j] = z > 0.f ? z : z * neg_slope;
                        }
                    }
                }
                if (fmode == 4) {
                    // same two roundings as conv -> bf16 -> activation -> bf16 (what the reference's bf16 pipeline and
                    // the stand-alone activation pass do), so fused and unfused layers are bit-identical
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float r = __bfloat162float(__float2bfloat16_rn(f[j]));
                        f[j] = r > 0.f ? r : r * neg_slope;
                    }
                }
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                if (tep) {
                    // (rows / columns outside the tensor are dropped by the bulk store itself)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        st_shared_v4(my + ((static_cast<uint32_t>(k) ^ sw) << 4), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2],
                                     pk[4 * k + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d(&p.omap[phase], sE + (ew * S + slot) * kSlabBytes, n0 + c, j0 + sub_w, i0 + sub_h,
                                     b0 + sub_b);
                        bulk_commit();
                    }
                    if (++slot == S) { slot = 0; xpar ^= 1; }
                } else if (valid) {
                    uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + off + c);
                    dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    if (full_chunk) {
                        dst[2] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
                        dst[3] = make_uint4(pk[12], pk[13], pk[14], pk[15]);
                    }
                }
                if (fmode == 1 || fmode == 2) {
                    // statistics of exactly what was stored (the bf16-rounded values), rows outside the tensor excluded
                    float s1[32];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        f[2 * j] = __uint_as_float(pk[j] << 16);
                        f[2 * j + 1] = __uint_as_float(pk[j] & 0xFFFF0000u);
                    }
                    if (!all_valid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = valid ? f[j] : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) s1[j] = f[j] * (fmode == 1 ? f[j] : xv[j]);
                    const float t0 = warp_colsum32(f, lane);
                    const float t1 = warp_colsum32(s1, lane);
                    if (lane < 16 || full_chunk) {
                        atomicAdd(gs0 + ch0 + lane, t0);
                        atomicAdd(gs1 + ch0 + lane, t1);
                    }
                }
            }
            // this warp is done reading the accumulator: hand it back to the MMA thread
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (tep && lane == 0) bulk_wait<0>();      // the slabs must outlive the stores that read them
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols);
    if (fmode == 1 || fmode == 2) {
        // one flush per CTA; a persistent CTA touches one or two N tiles, the rest of its table is still zero
        for (int i = threadIdx.x; i < p.fuse_groups * 2 * fC; i += blockDim.x) {
            const float t = s_sums[i];
            if (t != 0.f) atomicAdd(p.fuse_sums + i, t);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// wgrad-type kernel: both operands MN-major (pixels are the reduction dimension).
// Two smem rings: the P tile of a pixel block is loaded ONCE and reused by every tap of the CTA's tap group,
// the shifted Q tiles stream through their own ring.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// experiment switch: -DVG_WGRAD_SPIN=1 makes the producer / MMA threads of the weight-gradient kernel busy-poll
// (measured: no difference - the per-stage time of this kernel is not barrier-wake-up latency)
#ifndef VG_WGRAD_SPIN
#define VG_WGRAD_SPIN 0
#endif
#if VG_WGRAD_SPIN
#define WGRAD_WAIT mbar_wait_spin
#else
#define WGRAD_WAIT mbar_wait
#endif
// -DVG_WGRAD_TRACE=1 (make trace): CTA (0,0,0) accumulates the cycles its producer / issuer threads spend per phase
//   [0] producer: wait empty_b   [1] producer: expect_tx + Q TMA issue   [2] producer: wait empty_a   [3] P TMA issue
//   [4] issuer: wait full_b      [5] issuer: UMMA issue                  [6] issuer: commit + ring    [7] wait full_a
#ifndef VG_WGRAD_TRACE
#define VG_WGRAD_TRACE 0
#endif
// -DVG_DEBUG_WGRAD=1 (make debug-wgrad): WgradParams::debug_flags is honoured (profiling experiments that skip the
// epilogue / main loop / MMAs / TMA loads; results are wrong).  The default build carries none of those branches.
#ifndef VG_DEBUG_WGRAD
#define VG_DEBUG_WGRAD 0
#endif
#if VG_DEBUG_WGRAD
#define WGRAD_DBG(p, bit) (((p).debug_flags & (bit)) != 0)
#else
#define WGRAD_DBG(p, bit) false
#endif
#if VG_WGRAD_TRACE
__device__ unsigned long long g_wgrad_trace[8];
// counters live in registers and are flushed once per role (a global read-modify-write per phase would time itself)
#define TR_BEGIN()            \
    long long tr_t = clock64(); \
    long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define TR_ADD(i)                         \
    do {                                  \
        const long long tr_n = clock64(); \
        tr_acc[i] += tr_n - tr_t;         \
        tr_t = tr_n;                      \
    } while (0)
#define TR_FLUSH()                                                          \
    do {                                                                    \
        if (tr_on)                                                          \
            for (int tr_i = 0; tr_i < 8; ++tr_i) g_wgrad_trace[tr_i] += tr_acc[tr_i]; \
    } while (0)
#else
#define TR_BEGIN() (void)0
#define TR_ADD(i) (void)0
#define TR_FLUSH() (void)0
#endif
// PAIR = true: launched as clusters of two CTAs along grid.x holding consecutive M tiles of the same N tile; the pair
// issues M = 256 UMMAs (cta_group::2).  Each CTA loads its own P tile and HALF of the Q atoms of every stage, so the
// Q operand - the bulk of the L2 -> shared-memory traffic that bounds this kernel - is fetched once per pair.
template <bool PAIR>
__device__ __forceinline__ void wgrad_body(const WgradParams& p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    // (warp index through a shuffle: the compiler then KNOWS it is warp-uniform and keeps the role branches, and every
    // address / descriptor computed inside them, on the uniform datapath instead of "waterfall" R2UR loops)
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int kpix = p.tw * p.th * p.tb;
    const int p_row = p.p_atom_c * 2, q_row = p.q_atom_c * 2;       // bytes per pixel row inside one atom
    const int p_atom_bytes = kpix * p_row, q_atom_bytes = kpix * q_row;
    const int n_atoms = p.n_tile / p.q_atom_c;
    const int merge = p.merge > 1 ? p.merge : 1;
    const int a_stage = kpix * 256;                                  // room for all 128 M rows (128 ch x 2 B) per pixel
    const int tap_bytes = n_atoms * q_atom_bytes;                    // one tap's Q tile
    const int b_stage = merge * tap_bytes / (PAIR ? 2 : 1);          // a pair CTA holds half of the stage's Q atoms
    const int SA = p.stages_a, SB = p.stages_b;
    uint8_t* sA = smem;
    uint8_t* sB = smem + SA * a_stage;
    uint64_t* full_a = reinterpret_cast<uint64_t*>(sB + SB * b_stage);
    uint64_t* empty_a = full_a + kMaxStages;
    uint64_t* full_b = empty_a + kMaxStages;
    uint64_t* empty_b = full_b + kMaxStages;
    uint64_t* tmem_full = empty_b + kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    // plain: grid = (splits, tiles, tap groups).  pairs: grid = (tiles, splits, tap groups) with the cluster along x
    // (a CTA pair is two consecutive x ranks) and tile = n_tile_idx * m_tiles + m_tile, i.e. M tiles 2j and 2j+1
    const int split = PAIR ? blockIdx.y : blockIdx.x;
    const int tile_idx = PAIR ? blockIdx.x : blockIdx.y;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int m_tile = PAIR ? tile_idx % p.m_tiles : tile_idx / p.n_tiles;
    const int n_tile_idx = PAIR ? tile_idx / p.m_tiles : tile_idx % p.n_tiles;
    const int m0 = m_tile * p.m_atoms * p.p_atom_c, n0 = n_tile_idx * p.n_tile;
    const int tap0 = blockIdx.z * p.taps_per_cta;
    const int ntap = min(p.taps_per_cta, p.num_taps - tap0);
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    const int pt_begin = static_cast<int>(static_cast<long long>(total_tiles) * split / p.splits);
    int pt_end = static_cast<int>(static_cast<long long>(total_tiles) * (split + 1) / p.splits);
    if (WGRAD_DBG(p, 2)) pt_end = pt_begin;
    const uint32_t ncols = tmem_cols_for(p.taps_per_cta * p.n_tile);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.pmap);
        for (int v = 0; v < 4; ++v) tma_prefetch_desc(&p.qmap[v]);
        for (int s = 0; s < SA; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        if (PAIR) { tmem_alloc_pair(tmem_slot, ncols); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_slot, ncols); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();          // both CTAs' barriers exist before anything arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_enter();                           // (pdl.cuh: the prologue above ran under the previous kernel's tail)

    const bool leader = lane == 0;        // see igemm_fprop_kernel: all lanes run the loops, lane 0 issues
    if (warp == 0) {
        {
            int sa = 0, sb = 0;
            uint32_t par_a = 0, par_b = 0;
            int t0 = pt_begin;
            int tj = t0 % p.tiles_w;
            t0 /= p.tiles_w;
            int ti = t0 % p.tiles_h, tb_i = t0 / p.tiles_h;
            const bool tr_on = leader && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
            (void)tr_on;
            TR_BEGIN();
            for (int pt = pt_begin; pt < pt_end; ++pt) {
                const int j0 = tj * p.tw, i0 = ti * p.th, b0 = tb_i * p.tb;
                WGRAD_WAIT(&empty_a[sa], par_a ^ 1);
                TR_ADD(2);
                if (elect_one()) {     // (elect.sync: straight-line TMA issue, see igemm_fprop_kernel)
                    if (PAIR) {
                        // CTA 0's barrier counts the bytes of both CTAs; each CTA's data lands in its own shared memory
                        if (rank == 0) mbar_expect_tx(&full_a[sa], 2 * p.m_atoms * p_atom_bytes);
                        for (int a = 0; a < p.m_atoms; ++a)
                            tma_load_4d_pair(sA + sa * a_stage + a * p_atom_bytes, &p.pmap, &full_a[sa],
                                             m0 + a * p.p_atom_c, j0, i0, b0);
                    } else if (WGRAD_DBG(p, 8)) {
                        mbar_arrive(&full_a[sa]);
                    } else {
                        mbar_expect_tx(&full_a[sa], p.m_atoms * p_atom_bytes);
                        for (int a = 0; a < p.m_atoms; ++a)
                            tma_load_4d(sA + sa * a_stage + a * p_atom_bytes, &p.pmap, &full_a[sa], m0 + a * p.p_atom_c,
                                        j0, i0, b0);
                    }
                }
                __syncwarp();
                TR_ADD(3);
                for (int tl = 0; tl < ntap; tl += merge) {
                    const int cnt = min(merge, ntap - tl);
                    WGRAD_WAIT(&empty_b[sb], par_b ^ 1);
                    TR_ADD(0);
                    if (elect_one()) {
                        if (PAIR) {
                            // this CTA's half of the stage's Q atoms (atom index = tap_local * n_atoms + atom)
                            const int half = cnt * n_atoms / 2;
                            if (rank == 0) mbar_expect_tx(&full_b[sb], cnt * tap_bytes);
                            for (int h = 0; h < half; ++h) {
                                const int idx = static_cast<int>(rank) * half + h;
                                const int j = idx / n_atoms, a = idx - j * n_atoms;
                                const IgemmTap tap = p.taps[tap0 + tl + j];
                                tma_load_4d_pair(sB + sb * b_stage + h * q_atom_bytes, &p.qmap[tap.view], &full_b[sb],
                                                 n0 + a * p.q_atom_c, j0 + tap.dx, i0 + tap.dy, b0);
                            }
                        } else if (WGRAD_DBG(p, 8)) {
                            mbar_arrive(&full_b[sb]);
                        } else {
                            mbar_expect_tx(&full_b[sb], cnt * tap_bytes);
                            for (int j = 0; j < cnt; ++j) {
                                const IgemmTap tap = p.taps[tap0 + tl + j];
                                for (int a = 0; a < n_atoms; ++a)
                                    tma_load_4d(sB + sb * b_stage + j * tap_bytes + a * q_atom_bytes, &p.qmap[tap.view],
                                                &full_b[sb], n0 + a * p.q_atom_c, j0 + tap.dx, i0 + tap.dy, b0);
                            }
                        }
                    }
                    __syncwarp();
                    if (++sb == SB) { sb = 0; par_b ^= 1; }
                    TR_ADD(1);
                }
                if (++sa == SA) { sa = 0; par_a ^= 1; }
                if (++tj == p.tiles_w) { tj = 0; if (++ti == p.tiles_h) { ti = 0; ++tb_i; } }
            }
            TR_FLUSH();
        }
    } else if (warp == 1) {
        if (rank == 0) {
            const uint32_t idesc_full = make_idesc_bf16(PAIR ? 256 : 128, merge * p.n_tile, 1, 1);
            const int tail = ntap % merge;
            const uint32_t idesc_tail = make_idesc_bf16(PAIR ? 256 : 128, (tail ? tail : merge) * p.n_tile, 1, 1);
            const uint32_t p_layout = p.p_atom_c == 64 ? 2u : (p.p_atom_c == 32 ? 4u : 6u);
            const uint32_t q_layout = p.q_atom_c == 64 ? 2u : (p.q_atom_c == 32 ? 4u : 6u);
            const int ksteps = kpix / 16;
            // one UMMA consumes 16 pixel rows: two 8-row groups (SBO apart); MN atoms are LBO apart.  Descriptors of
            // ring slot 0 / k-step 0 are built once; per instruction only the start-address field (lo word) moves.
            const uint64_t a_desc0 = make_smem_desc(smem_u32(sA), p_atom_bytes, 8 * p_row, p_layout);
            const uint64_t b_desc0 = make_smem_desc(smem_u32(sB), q_atom_bytes, 8 * q_row, q_layout);
            const uint32_t a_hi = static_cast<uint32_t>(a_desc0 >> 32), b_hi = static_cast<uint32_t>(b_desc0 >> 32);
            const uint32_t a_lo0 = static_cast<uint32_t>(a_desc0), b_lo0 = static_cast<uint32_t>(b_desc0);
            const uint32_t a_step = a_stage >> 4, b_step = b_stage >> 4;
            const uint32_t a_k = (16 * p_row) >> 4, b_k = (16 * q_row) >> 4;
            int sa = 0, sb = 0;
            uint32_t par_a = 0, par_b = 0, a_lo = a_lo0, b_lo = b_lo0;
            const bool tr_on = leader && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
            (void)tr_on;
            TR_BEGIN();
            for (int pt = pt_begin; pt < pt_end; ++pt) {
                WGRAD_WAIT(&full_a[sa], par_a);
                TR_ADD(7);
                const uint32_t acc = pt != pt_begin;
                uint32_t d_tmem = tmem_base;
                for (int tl = 0; tl < ntap; tl += merge) {
                    const uint32_t idesc = (tl + merge <= ntap) ? idesc_full : idesc_tail;
                    WGRAD_WAIT(&full_b[sb], par_b);
#if VG_STAGE_FENCE
                    tc_fence_after();
#endif
                    TR_ADD(4);
                    if (!elect_one()) {
                        // (one elected lane issues; the other lanes follow the ring with it)
                    } else if (PAIR) {
                        for (int k = 0; k < ksteps; ++k)
                            umma_pair_bf16_lohi(d_tmem, a_lo + k * a_k, a_hi, b_lo + k * b_k, b_hi, idesc, acc | (k != 0));
                        umma_commit_pair(&empty_b[sb]);
                    } else if (WGRAD_DBG(p, 4)) {
                        mbar_arrive(&empty_b[sb]);
                    } else {
                        if (ksteps == 8) {
                            umma_bf16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, acc);
#pragma unroll
                            for (int k = 1; k < 8; ++k)
                                umma_bf16_lohi(d_tmem, a_lo + k * a_k, a_hi, b_lo + k * b_k, b_hi, idesc, 1);
                        } else if (ksteps == 4) {
                            umma_bf16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, acc);
                            umma_bf16_lohi(d_tmem, a_lo + a_k, a_hi, b_lo + b_k, b_hi, idesc, 1);
                            umma_bf16_lohi(d_tmem, a_lo + 2 * a_k, a_hi, b_lo + 2 * b_k, b_hi, idesc, 1);
                            umma_bf16_lohi(d_tmem, a_lo + 3 * a_k, a_hi, b_lo + 3 * b_k, b_hi, idesc, 1);
                        } else {
                            for (int k = 0; k < ksteps; ++k)
                                umma_bf16_lohi(d_tmem, a_lo + k * a_k, a_hi, b_lo + k * b_k, b_hi, idesc, acc | (k != 0));
                        }
                        TR_ADD(5);
                        umma_commit(&empty_b[sb]);
                    }
                    __syncwarp();
                    d_tmem += merge * p.n_tile;
                    b_lo += b_step;
                    if (++sb == SB) { sb = 0; par_b ^= 1; b_lo = b_lo0; }
                    TR_ADD(6);
                }
                if (elect_one()) {
                    if (PAIR) umma_commit_pair(&empty_a[sa]);
                    else if (WGRAD_DBG(p, 4)) mbar_arrive(&empty_a[sa]);
                    else umma_commit(&empty_a[sa]);
                }
                __syncwarp();
                a_lo += a_step;
                if (++sa == SA) { sa = 0; par_a ^= 1; a_lo = a_lo0; }
            }
            if (elect_one()) {
                if (PAIR) umma_commit_pair(tmem_full);
                else umma_commit(tmem_full);
            }
            TR_FLUSH();
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int m = m0 + row;
        const bool valid = (row < p.m_atoms * p.p_atom_c) && (m < p.m_valid);
        const bool has_work = (pt_end > pt_begin) && !WGRAD_DBG(p, 1);
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        if (p.atomic_split) {
            // split layers: dw[m][n][4 taps] += this split's tile, one 16-byte reduction per (m, n, 4 taps)
            for (int g4 = 0; g4 < ntap; g4 += 4) {
                float* dst = p.dw + static_cast<long long>(m) * p.s_m + (tap0 + g4);
                for (int c = 0; c < p.n_tile; c += 16) {
                    uint32_t v0[16], v1[16], v2[16], v3[16];
                    tmem_ld_32x16(taddr + (g4 + 0) * p.n_tile + c, v0);
                    tmem_ld_32x16(taddr + (g4 + 1) * p.n_tile + c, v1);
                    tmem_ld_32x16(taddr + (g4 + 2) * p.n_tile + c, v2);
                    tmem_ld_32x16(taddr + (g4 + 3) * p.n_tile + c, v3);
                    tmem_ld_wait();
                    if (valid && has_work) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int n = n0 + c + j;
                            if (n < p.n_valid)
                                red_add_v4(dst + static_cast<long long>(n) * p.s_n, __uint_as_float(v0[j]),
                                           __uint_as_float(v1[j]), __uint_as_float(v2[j]), __uint_as_float(v3[j]));
                        }
                    }
                }
            }
        } else if (p.splits > 1) {
            // partial tile -> workspace: [cta][tap_local][row][n_tile]; each thread writes its row contiguously
            const long long y_canon = static_cast<long long>(m_tile) * p.n_tiles + n_tile_idx;
            const long long cta = (static_cast<long long>(blockIdx.z) * (p.m_tiles * p.n_tiles) + y_canon) * p.splits + split;
            float* base = p.partial + cta * static_cast<long long>(p.taps_per_cta) * 128 * p.n_tile;
            for (int tl = 0; tl < ntap; ++tl) {
                float4* dst = reinterpret_cast<float4*>(base + (static_cast<long long>(tl) * 128 + row) * p.n_tile);
                for (int c = 0; c < p.n_tile; c += 16) {
                    uint32_t v[16];
                    tmem_ld_32x16(taddr + tl * p.n_tile + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        dst[c / 4 + j] = has_work ? make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        } else {
            // exclusive ownership: dw[m][n][tap] += acc, no atomics
            if (p.vec4_taps) {
                for (int g4 = 0; g4 < ntap; g4 += 4) {
                    float* dst = p.dw + static_cast<long long>(m) * p.s_m + (tap0 + g4);
                    for (int c = 0; c < p.n_tile; c += 16) {
                        uint32_t v0[16], v1[16], v2[16], v3[16];
                        tmem_ld_32x16(taddr + (g4 + 0) * p.n_tile + c, v0);
                        tmem_ld_32x16(taddr + (g4 + 1) * p.n_tile + c, v1);
                        tmem_ld_32x16(taddr + (g4 + 2) * p.n_tile + c, v2);
                        tmem_ld_32x16(taddr + (g4 + 3) * p.n_tile + c, v3);
                        tmem_ld_wait();
                        if (valid && has_work) {
                            // all 16 read-modify-writes of this chunk in flight together: loads first, stores after
                            float4 old[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int n = min(n0 + c + j, p.n_valid - 1);
                                old[j] = (WGRAD_DBG(p, 16) || !p.accumulate)
                                             ? make_float4(0.f, 0.f, 0.f, 0.f)
                                             : __ldcg(reinterpret_cast<const float4*>(dst + static_cast<long long>(n) * p.s_n));
                            }
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int n = n0 + c + j;
                                if (n < p.n_valid) {
                                    float4 o = old[j];
                                    o.x += __uint_as_float(v0[j]);
                                    o.y += __uint_as_float(v1[j]);
                                    o.z += __uint_as_float(v2[j]);
                                    o.w += __uint_as_float(v3[j]);
                                    __stcg(reinterpret_cast<float4*>(dst + static_cast<long long>(n) * p.s_n), o);
                                }
                            }
                        }
                    }
                }
            } else {
                for (int tl = 0; tl < ntap; ++tl) {
                    const int tap_id = p.taps[tap0 + tl].tap_id;
                    float* dst = p.dw + static_cast<long long>(m) * p.s_m + static_cast<long long>(tap_id) * p.s_tap;
                    for (int c = 0; c < p.n_tile; c += 16) {
                        uint32_t v[16];
                        tmem_ld_32x16(taddr + tl * p.n_tile + c, v);
                        tmem_ld_wait();
                        if (valid && has_work) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int n = n0 + c + j;
                                if (n < p.n_valid) {
                                    float* q = dst + static_cast<long long>(n) * p.s_n;
                                    *q = (p.accumulate ? *q : 0.f) + __uint_as_float(v[j]);
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) {
        cluster_sync_all();            // the peer's shared memory / barriers stay alive until both CTAs are done
        if (warp == 1) tmem_dealloc_pair(tmem_base, ncols);
    } else {
        if (warp == 1) tmem_dealloc(tmem_base, ncols);
    }
}

__global__ void __launch_bounds__(kIgemmThreads) igemm_wgrad_kernel(const __grid_constant__ WgradParams p) {
    wgrad_body<false>(p);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kIgemmThreads)
    igemm_wgrad_pair_kernel(const __grid_constant__ WgradParams p) {
    wgrad_body<true>(p);
}

// dw[m][n][tap] += sum over splits of partial[cta(split, mn, group)][tap_local][row][col]
// VEC = 4: one thread owns the 4 consecutive taps dw[m][n][4t..4t+3]; VEC = 1: one tap.
// One thread per output unit, n fastest (coalesced partial reads).  When there are few outputs but many splits the
// split range is cut into gridDim.y chunks whose sums are combined with fp32 atomics (CHUNKED = true).
template <int VEC, bool CHUNKED>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const WgradParams p) {
    pdl_enter();
    const int tap_units = p.num_taps / VEC;
    const long long total = static_cast<long long>(p.m_valid) * p.n_valid * tap_units;
    const int rows_per_tile = p.m_atoms * p.p_atom_c;
    const long long plane = static_cast<long long>(128) * p.n_tile;
    const long long cta_stride = static_cast<long long>(p.taps_per_cta) * plane;
    const int s_begin = CHUNKED ? static_cast<int>(static_cast<long long>(p.splits) * blockIdx.y / gridDim.y) : 0;
    const int s_end = CHUNKED ? static_cast<int>(static_cast<long long>(p.splits) * (blockIdx.y + 1) / gridDim.y) : p.splits;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        // i enumerates (m, tap unit, n) with n fastest so that the partial reads are coalesced
        const int n = static_cast<int>(i % p.n_valid);
        const long long r = i / p.n_valid;
        const int tap = static_cast<int>(r % tap_units) * VEC;
        const int m = static_cast<int>(r / tap_units);
        const int m_tile = m / rows_per_tile, row = m - m_tile * rows_per_tile;
        const int n_tile_idx = n / p.n_tile, col = n - n_tile_idx * p.n_tile;
        const int group = tap / p.taps_per_cta, tl = tap - group * p.taps_per_cta;
        const long long y = static_cast<long long>(m_tile) * p.n_tiles + n_tile_idx;
        const float* src = p.partial + ((static_cast<long long>(group) * (p.m_tiles * p.n_tiles) + y) * p.splits) * cta_stride +
                           static_cast<long long>(tl) * plane + static_cast<long long>(row) * p.n_tile + col;
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        for (int s = s_begin; s < s_end; ++s) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] += __ldcs(src + s * cta_stride + v * plane);
        }
        float* dst = p.dw + static_cast<long long>(m) * p.s_m + static_cast<long long>(n) * p.s_n +
                     static_cast<long long>(tap) * p.s_tap;
        if (CHUNKED) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) atomicAdd(dst + v, acc[v]);
        } else if (VEC == 4) {
            float4 o = p.accumulate ? *reinterpret_cast<float4*>(dst) : make_float4(0.f, 0.f, 0.f, 0.f);
            o.x += acc[0]; o.y += acc[1]; o.z += acc[2]; o.w += acc[3];
            *reinterpret_cast<float4*>(dst) = o;
        } else {
            dst[0] = (p.accumulate ? dst[0] : 0.f) + acc[0];
        }
    }
}

__global__ void splitk_finish_kernel(const float* __restrict__ acc, const float* __restrict__ bias, void* out,
                                     int out_fp32, size_t n, int C) {
    pdl_enter();
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float v = acc[i];
        if (bias != nullptr) v += bias[i % C];
        if (out_fp32) static_cast<float*>(out)[i] = v;
        else static_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int smem_bytes_for(int stages, int stage_bytes) { return stages * stage_bytes + 1024 + kBarrierBytes; }

// dynamic shared memory of an fprop-type launch: operand ring, barriers, fused-epilogue tables
int igemm_total_smem(const IgemmParams& p) {
    const int tps = p.tps > 1 ? p.tps : 1;
    const int stage = p.halo ? p.halo_stage_bytes + tps * p.n_tile * p.kchunk * 2 : tps * (128 + p.n_tile) * p.kchunk * 2;
    return p.stages * stage + 1024 + igemm_epilogue_smem_bytes(p) + kBarrierBytes + igemm_fuse_smem_bytes(p);
}

int launch_igemm(const IgemmParams& p, cudaStream_t stream) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        for (auto fn : {igemm_fprop_kernel<0>, igemm_fprop_kernel<1>, igemm_fprop_kernel<2>, igemm_fprop_kernel<3>,
                        igemm_fprop_kernel<4>, igemm_fprop_kernel<5>, igemm_fprop_kernel<5, true>,
                        igemm_fprop_kernel<0, true>, igemm_fprop_kernel<1, true>,
                        igemm_fprop_kernel<2, true>, igemm_fprop_kernel<3, true>, igemm_fprop_kernel<4, true>}) {
            const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) attr_err = e;
        }
    });
    if (attr_err != cudaSuccess) return static_cast<int>(attr_err);
    const int smem = igemm_total_smem(p);
    const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
    const long long items = static_cast<long long>(p.tiles_w) * p.tiles_h * p.tiles_b * p.n_tiles * p.num_phases * ksplit;
    const int per_sm = smem <= 112 * 1024 ? 2 : 1;
    const int threads = p.ew8 ? kIgemmMaxThreads : kIgemmThreads;
    dim3 grid(static_cast<unsigned>(std::min<long long>(items, 148LL * per_sm)));
    const size_t out_elems = static_cast<size_t>(p.out_B) * p.out_H * p.out_W * p.out_C;
    if (ksplit > 1) {
        cudaError_t e = cudaMemsetAsync(p.splitk_acc, 0, out_elems * sizeof(float), stream);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    if (p.halo) {
        switch (p.fuse_mode) {
            case 1: launch_k(igemm_fprop_kernel<1, true>, dim3(grid), dim3(threads), smem, stream, p); break;
            case 2: launch_k(igemm_fprop_kernel<2, true>, dim3(grid), dim3(threads), smem, stream, p); break;
            case 3: launch_k(igemm_fprop_kernel<3, true>, dim3(grid), dim3(threads), smem, stream, p); break;
            case 4: launch_k(igemm_fprop_kernel<4, true>, dim3(grid), dim3(threads), smem, stream, p); break;
            case 5: launch_k(igemm_fprop_kernel<5, true>, dim3(grid), dim3(threads), smem, stream, p); break;
            default: launch_k(igemm_fprop_kernel<0, true>, dim3(grid), dim3(threads), smem, stream, p); break;
        }
    } else
    switch (p.fuse_mode) {
        case 1: launch_k(igemm_fprop_kernel<1>, dim3(grid), dim3(threads), smem, stream, p); break;
        case 2: launch_k(igemm_fprop_kernel<2>, dim3(grid), dim3(threads), smem, stream, p); break;
        case 3: launch_k(igemm_fprop_kernel<3>, dim3(grid), dim3(threads), smem, stream, p); break;
        case 4: launch_k(igemm_fprop_kernel<4>, dim3(grid), dim3(threads), smem, stream, p); break;
        case 5: launch_k(igemm_fprop_kernel<5>, dim3(grid), dim3(threads), smem, stream, p); break;
        default: launch_k(igemm_fprop_kernel<0>, dim3(grid), dim3(threads), smem, stream, p); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
    if (ksplit > 1) {
        const int blocks = static_cast<int>(std::min<size_t>((out_elems + 255) / 256, 148 * 8));
        launch_k(splitk_finish_kernel, dim3(blocks), dim3(256), 0, stream, p.splitk_acc, p.bias, p.out, p.out_fp32, out_elems, p.out_C);
        e = cudaGetLastError();
    }
    return static_cast<int>(e);
}

int launch_wgrad(const WgradParams& p, cudaStream_t stream) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(igemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        const cudaError_t e2 =
            cudaFuncSetAttribute(igemm_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (attr_err == cudaSuccess) attr_err = e2;
    });
    if (attr_err != cudaSuccess) return static_cast<int>(attr_err);
    const int kpix = p.tw * p.th * p.tb;
    const int smem = p.stages_a * kpix * 256 +
                     p.stages_b * (p.merge > 1 ? p.merge : 1) * p.n_tile * kpix * 2 / (p.pair ? 2 : 1) + 1024 + kBarrierBytes;
    const int tap_groups = (p.num_taps + p.taps_per_cta - 1) / p.taps_per_cta;
    dim3 grid(p.splits, p.m_tiles * p.n_tiles, tap_groups);
    cudaError_t e;
    if (p.pair) {
        const dim3 pgrid(p.m_tiles * p.n_tiles, p.splits, tap_groups);          // cluster (2, 1, 1) from the kernel attribute
        launch_k(igemm_wgrad_pair_kernel, dim3(pgrid), dim3(kIgemmThreads), smem, stream, p);
        e = cudaGetLastError();
    } else {
        launch_k(igemm_wgrad_kernel, dim3(grid), dim3(kIgemmThreads), smem, stream, p);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess || p.splits <= 1 || p.atomic_split) return static_cast<int>(e);
    const long long total = static_cast<long long>(p.m_valid) * p.n_valid * p.num_taps / (p.vec4_taps ? 4 : 1);
    const bool chunked = p.splits >= 16 && total <= 32768;   // few outputs, many splits
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
    const dim3 rgrid(blocks, chunked ? std::min(p.splits / 4, 32) : 1);
    if (p.vec4_taps) {
        if (chunked) launch_k(wgrad_reduce_kernel<4, true>, dim3(rgrid), dim3(256), 0, stream, p);
        else launch_k(wgrad_reduce_kernel<4, false>, dim3(rgrid), dim3(256), 0, stream, p);
    } else {
        if (chunked) launch_k(wgrad_reduce_kernel<1, true>, dim3(rgrid), dim3(256), 0, stream, p);
        else launch_k(wgrad_reduce_kernel<1, false>, dim3(rgrid), dim3(256), 0, stream, p);
    }
    return static_cast<int>(cudaGetLastError());
}

size_t wgrad_partial_bytes(const WgradParams& p) {
    const int tap_groups = (p.num_taps + p.taps_per_cta - 1) / p.taps_per_cta;
    return static_cast<size_t>(p.splits) * p.m_tiles * p.n_tiles * tap_groups * p.taps_per_cta * 128 * p.n_tile *
           sizeof(float);
}

#if VG_WGRAD_TRACE
extern "C" int vg_debug_wgrad_trace(unsigned long long* out8, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, g_wgrad_trace, sizeof(unsigned long long) * 8);
    if (reset) {
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(g_wgrad_trace, z, sizeof(z));
    }
    return 0;
}
#endif

int igemm_smem_bytes(int stages, int stage_bytes) { return smem_bytes_for(stages, stage_bytes); }

// cuTensorMapEncodeTiled is a driver entry point; fetch it through the runtime so the library links without
// libcuda (the build container has no driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return -1;
    cuuint64_t gdim[5], gstride[5];
    cuuint32_t gbox[5], estride[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        gbox[i] = box[i];
        estride[i] = 1;
        if (i > 0) gstride[i - 1] = strides_elems[i] * 2;  // bytes
    }
    const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                        : CU_TENSOR_MAP_SWIZZLE_32B;
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                          gdim, gstride, gbox, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

}  // namespace vg
