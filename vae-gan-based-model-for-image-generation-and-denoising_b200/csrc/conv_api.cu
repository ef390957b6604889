// C-ABI for the three convolution contractions (down / up / wgrad): geometry -> tap tables + TMA tensor maps
// -> tcgen05 implicit-GEMM launch (bf16), or the fp32 CUDA-core implicit GEMM (conv_simt.cu).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include <cuda_bf16.h>

#include "common.cuh"
#include "pdl.cuh"
#include "conv_simt.cuh"
#include "gemv.cuh"
#include "igemm_umma.cuh"

namespace vg {

// Experiment switches (environment variables): read ONCE, at the first launch, never on the
// launch path.  All unset in the product path.
struct Switches {
    int halo;            // VG_HALO: 0 off, 1 every eligible launch, n >= 16 only N tiles <= n
    int wgrad_kpix;      // VG_WGRAD_KPIX: pixels per stage of the weight-gradient kernel (default 128)
    int wgrad_sa;        // VG_WGRAD_SA: depth of its P ring (0 = default)
    int wgrad_pair;      // VG_WGRAD2: CTA pairs
    int wgrad_atomic;    // VG_WGRAD_ATOMIC: split layers reduce with vector atomics (1) or partial tiles + reduce (0)
    int debug_wgrad;     // VG_DEBUG_WGRAD: only honoured by -DVG_DEBUG_WGRAD=1 builds
    int tep;             // VG_TEP: 0 = per-lane global loads / stores in the epilogues instead of the TMA epilogue
    int ew8;             // VG_EW8: 0 = four epilogue warps even for launches that own a whole SM
    int wide_from;       // VG_WIDE_FROM: N tiles >= this own a whole SM (default 256; 128 measured 15-25 % slower on the N = 128 layers)
    int xtma_wide;       // VG_XTMA_WIDE: N = 256 tiles of modes 2 / 3 also fetch the saved tensor by TMA
};
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
static const Switches& switches() {
    static const Switches s = {
        getenv("VG_HALO") ? std::max(1, atoi(getenv("VG_HALO"))) : 0,
        env_int("VG_WGRAD_KPIX", 128),
        env_int("VG_WGRAD_SA", 0),
        getenv("VG_WGRAD2") != nullptr ? 1 : 0,
        env_int("VG_WGRAD_ATOMIC", 1),
        env_int("VG_DEBUG_WGRAD", 0),
        env_int("VG_TEP", 1),
        env_int("VG_EW8", 1),
        env_int("VG_WIDE_FROM", 256),
        env_int("VG_XTMA_WIDE", 0),
    };
    return s;
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }
static int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
static int floor_mod(int a, int m) { return ((a % m) + m) % m; }

// Box (tw, th, tb) with tw*th*tb == rows covering a (w, h, b) grid with as little waste as possible.
static void pick_box(int rows, int w, int h, int* tw, int* th, int* tb) {
    *tw = std::min(pow2_ceil(w), rows);
    *th = std::min(pow2_ceil(h), rows / *tw);
    *tb = rows / (*tw * *th);
}

static int pick_kchunk(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : (c % 16 == 0 ? 16 : 0)); }

static int pick_n_tile(int n) {
    for (int t = 256; t >= 16; t -= 16)
        if (n % t == 0) return t;
    return 0;
}

// Narrow-channel layers (C == 16 or 32, one channel chunk): put several taps into one pipeline stage.
static int pick_tps(int kchunk, int c_chunks, int taps) {
    if (c_chunks != 1 || kchunk >= 64) return 1;
    const int want = kchunk == 16 ? 4 : 2;
    for (int t = want; t > 1; --t)
        if (taps % t == 0) return t;
    return 1;
}

// N tiles below switches().wide_from run two CTAs per SM (two MMA-issuing threads, 4 + 4 epilogue warps); wider ones
// own the SM (one issuer is enough at >= 64 tensor cycles per instruction) and get eight epilogue warps.
static bool owns_sm(int n_tile) { return n_tile >= switches().wide_from; }

static int pick_stages(int stage_bytes, int n_tile, int iters = 1 << 30, int extra_smem = 0) {
    // The kernel is persistent (the ring runs across tiles), so the depth is purely a bandwidth / latency question.
    // `extra_smem`: epilogue slabs + fused-epilogue tables (barriers and alignment slack are budgeted here).
    (void)iters;
    const int budget = (owns_sm(n_tile) ? 225 : 112) * 1024 - 3072 - extra_smem;
    return std::max(2, std::min(8, budget / stage_bytes));
}

// NHWC view of `base` [B][H][W][C] decimated by `s` starting at (vy, vx); dims innermost first.
static int make_view(CUtensorMap* m, const void* base, int B, int H, int W, int C, int s, int vy, int vx, int box_c,
                     int tw, int th, int tb, int swizzle) {
    const int Wv = ceil_div(W - vx, s), Hv = ceil_div(H - vy, s);
    if (Wv <= 0 || Hv <= 0) {  // empty parity plane: never addressed in-bounds; keep a valid 1x1 window
        const uint64_t dims[4] = {(uint64_t)C, 1, 1, (uint64_t)B};
        const uint64_t strides[4] = {1, (uint64_t)C, (uint64_t)C, (uint64_t)H * W * C};
        const uint32_t box[4] = {(uint32_t)box_c, (uint32_t)tw, (uint32_t)th, (uint32_t)tb};
        return make_tmap_bf16(m, base, 4, dims, strides, box, swizzle);
    }
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wv, (uint64_t)Hv, (uint64_t)B};
    const uint64_t strides[4] = {1, (uint64_t)s * C, (uint64_t)s * W * C, (uint64_t)H * W * C};
    const uint32_t box[4] = {(uint32_t)box_c, (uint32_t)tw, (uint32_t)th, (uint32_t)tb};
    const char* p = static_cast<const char*>(base) + (static_cast<size_t>(vy) * W + vx) * C * 2;
    return make_tmap_bf16(m, p, 4, dims, strides, box, swizzle);
}

static int bc_valid(const VgConvGeom* g) { return g->big_c_valid > 0 ? g->big_c_valid : g->big_c; }

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int check_geom(const VgConvGeom* g) {
    if (g == nullptr) return fail(VG_ERR_ARG, "null geometry");
    if (g->batch <= 0 || g->big_h <= 0 || g->big_w <= 0 || g->big_c <= 0 || g->small_h <= 0 || g->small_w <= 0 ||
        g->small_c <= 0 || g->kernel <= 0 || g->stride <= 0 || g->pad < 0)
        return fail(VG_ERR_SHAPE, "non-positive extent in conv geometry");
    const int eh = (g->big_h + 2 * g->pad - g->kernel) / g->stride + 1;
    const int ew = (g->big_w + 2 * g->pad - g->kernel) / g->stride + 1;
    if (g->big_h + 2 * g->pad < g->kernel || g->big_w + 2 * g->pad < g->kernel)
        return fail(VG_ERR_SHAPE, "Kernel size can't be greater than actual input size");
    if (g->big_c_valid < 0 || g->big_c_valid > g->big_c) return fail(VG_ERR_SHAPE, "big_c_valid out of range");
    if (eh != g->small_h || ew != g->small_w)
        return fail(VG_ERR_SHAPE, "small extent %dx%d inconsistent with big %dx%d k%d s%d p%d (expect %dx%d)",
                    g->small_h, g->small_w, g->big_h, g->big_w, g->kernel, g->stride, g->pad, eh, ew);
    return VG_OK;
}

bool umma_down_ok(const VgConvGeom* g) {
    return (g->stride == 1 || g->stride == 2) && g->kernel * g->kernel <= 64 && pick_kchunk(g->big_c) != 0 &&
           pick_n_tile(g->small_c) != 0;
}
bool umma_up_ok(const VgConvGeom* g) {
    const bool dense = g->small_h == 1 && g->small_w == 1 && g->stride == 1 && g->pad == 0;
    if (pick_kchunk(g->small_c) == 0) return false;
    if (dense) return pick_n_tile(g->big_c) != 0;
    if (g->stride == 2 && (g->kernel % 2) != 0) return false;
    return (g->stride == 1 || g->stride == 2) && g->kernel * g->kernel <= 64 && pick_n_tile(g->big_c) != 0;
}
static int wgrad_n_tile(int big_c) {
    for (int t : {128, 64, 32, 16})
        if (big_c % t == 0) return t;
    return 0;
}
bool umma_wgrad_ok(const VgConvGeom* g) {
    return (g->stride == 1 || g->stride == 2) && g->kernel * g->kernel <= 16 && g->small_c % 16 == 0 &&
           wgrad_n_tile(g->big_c) != 0;
}

// ---------------------------------------------------------------------------------------------- fused epilogues
// Tiling decisions shared by the launchers and by vg_conv_epilogue_supported.
// Few work items (small late layers): narrower N tiles put more SMs to work; the UMMA time per item falls with N.
// Only well below one wave: 128 items of N = 256 on 148 SMs beat 256 items of N = 128 (47 vs 60 us on the
// 512->1024 4x4 layer) - wide tiles need fewer issue slots and fewer fill bytes per FLOP.
static int shrink_n_tile(int n_total, int n_tile, long long m_items) {
    while (m_items * (n_total / n_tile) < 96 && n_tile % 64 == 0 && n_total % (n_tile / 2) == 0) n_tile /= 2;
    return n_tile;
}

static void down_tiling(const VgConvGeom* g, int* tw, int* th, int* tb, int* n_tile) {
    pick_box(128, g->small_w, g->small_h, tw, th, tb);
    const long long m_items = static_cast<long long>(ceil_div(g->small_w, *tw)) * ceil_div(g->small_h, *th) *
                              ceil_div(g->batch, *tb);
    *n_tile = shrink_n_tile(g->small_c, pick_n_tile(g->small_c), m_items);
}
static void up_tiling(const VgConvGeom* g, int* tw, int* th, int* tb, int* n_tile, int* n_total) {
    const bool dense = g->small_h == 1 && g->small_w == 1 && g->stride == 1 && g->pad == 0;
    *n_total = dense ? g->kernel * g->kernel * g->big_c : g->big_c;
    const int grid_h = dense ? 1 : ceil_div(g->big_h, g->stride), grid_w = dense ? 1 : ceil_div(g->big_w, g->stride);
    pick_box(128, grid_w, grid_h, tw, th, tb);
    const long long m_items = static_cast<long long>(ceil_div(grid_w, *tw)) * ceil_div(grid_h, *th) *
                              ceil_div(g->batch, *tb) * (dense ? 1 : g->stride * g->stride);
    *n_tile = shrink_n_tile(*n_total, pick_n_tile(*n_total), m_items);
}

// 0 when `ep` can ride the tensor-core launch of this contraction, else a negative VG_ERR_* (message set).
static int fuse_check(const VgConvGeom* g, bool up, const VgEpilogue* ep) {
    if (ep == nullptr || ep->mode == VG_EPI_NONE) return VG_OK;
    if (up && is_gemv(g) && gemv_up_fused_ok(g, ep) && aligned16(ep->x)) return VG_OK;
    if (is_gemv(g) || !(up ? umma_up_ok(g) : umma_down_ok(g)))
        return fail(VG_ERR_SHAPE, "fused epilogue: this geometry does not run on the tensor-core path");
    int tw, th, tb, n_tile, n_total = g->small_c;
    if (up) up_tiling(g, &tw, &th, &tb, &n_tile, &n_total); else down_tiling(g, &tw, &th, &tb, &n_tile);
    if (ep->mode < VG_EPI_BN_STATS || ep->mode > VG_EPI_AFFINE_ACT_FWD) return fail(VG_ERR_ARG, "fused epilogue: bad mode %d", ep->mode);
    if (ep->mode == VG_EPI_AFFINE_ACT_FWD) {
        const int C = ep->channels;
        if (ep->stats == nullptr) return fail(VG_ERR_ARG, "fused epilogue: null scale / shift table");
        if (ep->act != VG_ACT_NONE && ep->act != VG_ACT_RELU && ep->act != VG_ACT_LEAKY)
            return fail(VG_ERR_SHAPE, "fused epilogue: only ReLU / LeakyReLU / identity ride the forward epilogue");
        if (ep->groups > 1) return fail(VG_ERR_SHAPE, "fused epilogue: the affine form has one parameter group");
        if (C <= 0 || C % 32 != 0 || n_tile % 32 != 0 || n_total % C != 0)
            return fail(VG_ERR_SHAPE, "fused epilogue: %d channels / N tile %d not multiples of 32", C, n_tile);
        if (static_cast<long long>(C) * 24 > 48 * 1024)
            return fail(VG_ERR_SHAPE, "fused epilogue: %d-channel table exceeds the shared-memory budget", C);
        return VG_OK;
    }
    if (ep->mode == VG_EPI_ACT_FWD) {
        if (ep->act != VG_ACT_RELU && ep->act != VG_ACT_LEAKY)
            return fail(VG_ERR_SHAPE, "fused epilogue: only ReLU / LeakyReLU ride the forward epilogue");
        return VG_OK;
    }
    if (ep->mode != VG_EPI_BN_STATS) {
        if (ep->x == nullptr) return fail(VG_ERR_ARG, "fused epilogue: saved conv output missing");
        if (ep->act != VG_ACT_NONE && ep->act != VG_ACT_RELU && ep->act != VG_ACT_LEAKY)
            return fail(VG_ERR_SHAPE, "fused epilogue: only ReLU / LeakyReLU / identity derivatives");
        if (!aligned16(ep->x)) return fail(VG_ERR_ALIGN, "fused epilogue: 16-byte alignment");
    }
    if (ep->mode == VG_EPI_ACT_BWD) return VG_OK;
    const int C = ep->channels, G = ep->groups;
    if (ep->sums == nullptr || (ep->mode == VG_EPI_BN_BWD && ep->stats == nullptr))
        return fail(VG_ERR_ARG, "fused epilogue: null statistics buffer");
    if (C <= 0 || C % 32 != 0 || n_tile % 32 != 0 || n_total % C != 0)
        return fail(VG_ERR_SHAPE, "fused epilogue: %d channels / N tile %d not multiples of 32", C, n_tile);
    if (G < 1 || g->batch % G != 0 || (G > 1 && (g->batch / G) % tb != 0))
        return fail(VG_ERR_SHAPE, "fused epilogue: %d statistics groups do not align with %d-image tiles", G, tb);
    if (static_cast<long long>(G) * C * 24 > 48 * 1024)
        return fail(VG_ERR_SHAPE, "fused epilogue: %d x %d channel table exceeds the shared-memory budget", G, C);
    return VG_OK;
}

static void fuse_fill(IgemmParams& p, const VgConvGeom* g, const VgEpilogue* ep) {
    if (ep == nullptr || ep->mode == VG_EPI_NONE) return;
    p.fuse_mode = ep->mode;
    p.fuse_groups = ep->groups > 0 ? ep->groups : 1;
    p.fuse_group_batch = g->batch / p.fuse_groups;
    p.fuse_c = ep->channels;
    p.fuse_sums = ep->sums;
    p.fuse_x = ep->x;
    p.fuse_stats = ep->stats;
    p.fuse_act = ep->act;
    p.fuse_slope = ep->slope;
}

// TMA epilogue plan (IgemmParams::tep / ew8 / ep_slots + the output-side tensor maps).  `out` / `x`: the output tensor
// and (modes 2 / 3) the saved tensor of the same NHWC shape [B][H][W][C]; strided phases (`up`, stride s) address
// their (row, column) parity views.  Call after n_tile, the tile box, the phases and ksplit are settled.
static void plan_epilogue_flags(IgemmParams& p, const void* out, int C, int out_f32) {
    p.ew8 = owns_sm(p.n_tile) && switches().ew8 != 0 && (p.n_tile / 32) % 2 == 0 && p.n_tile % 32 == 0;
    const bool mode23 = p.fuse_mode == VG_EPI_BN_BWD || p.fuse_mode == VG_EPI_ACT_BWD;
    p.tep = switches().tep != 0 && !out_f32 && p.ksplit <= 1 && p.n_tile % 32 == 0 && C % 8 == 0 && aligned16(out) &&
            (!mode23 || aligned16(p.fuse_x));
    // slabs per epilogue warp: two for plain stores; four when the saved tensor arrives by TMA two chunks ahead.
    // Launches that own the SM have eight epilogue warps and 48 KB stages: one slab each (the previous chunk's store
    // has long read it when a warp comes back two chunks later) keeps the fourth stage; at N = 256 the saved tensor
    // then comes per lane as before (VG_XTMA_WIDE=1: three slabs, TMA).
    if (mode23) p.ep_slots = (p.n_tile > 128 && !switches().xtma_wide) ? 1 : (p.ew8 ? 3 : 4);
    else p.ep_slots = p.ew8 ? 1 : 2;
}
static int plan_epilogue_maps(IgemmParams& p, const void* out, int B, int H, int W, int C) {
    const bool mode23 = p.fuse_mode == VG_EPI_BN_BWD || p.fuse_mode == VG_EPI_ACT_BWD;
    if (!p.tep) return 0;
    // the 32 consecutive tile rows of one epilogue warp form a (bw, bh, bb) box of the (tw, th, tb) tile
    const int bw = std::min(p.tw, 32), bh = std::min(p.th, 32 / bw), bb = 32 / (bw * bh);
    for (int ph = 0; ph < p.num_phases; ++ph) {
        int rc = make_view(&p.omap[ph], out, B, H, W, C, p.osy, p.ph_ay[ph], p.ph_ax[ph], 32, bw, bh, bb, 64);
        if (rc == 0 && mode23)
            rc = make_view(&p.xmap[ph], p.fuse_x, B, H, W, C, p.osy, p.ph_ay[ph], p.ph_ax[ph], 32, bw, bh, bb, 64);
        if (rc != 0) return rc;
    }
    return 0;
}
static int epilogue_extra_smem(const IgemmParams& p) { return igemm_epilogue_smem_bytes(p) + igemm_fuse_smem_bytes(p); }

// ---------------------------------------------------------------------------------------------- tap lists
// "down" (Conv2d forward / ConvTranspose2d dgrad): one phase, k*k taps in packed-weight order; tap (ky, kx) reads
// the (row parity, column parity) view of the big tensor that contains input row  s*i + ky - pad.
static void down_fill_taps(IgemmParams& p, int k, int s, int pad, int small_c) {
    for (int ky = 0; ky < k; ++ky)
        for (int kx = 0; kx < k; ++kx) {
            IgemmTap& t = p.taps[ky * k + kx];
            const int vy = floor_mod(ky - pad, s), vx = floor_mod(kx - pad, s);
            t.view = static_cast<int16_t>(vy * s + vx);
            t.dy = static_cast<int16_t>((ky - pad - vy) / s);
            t.dx = static_cast<int16_t>((kx - pad - vx) / s);
            t.tap_id = static_cast<int16_t>(ky * k + kx);
            t.brow = (ky * k + kx) * small_c;
        }
}

// "up" (ConvTranspose2d forward / Conv2d dgrad): s*s output phases (ay, ax); phase row parity ay takes the ky with
// (ay + pad - ky) divisible by s, from small row i + (ay + pad - ky) / s.  Sets num_phases, taps_per_phase, ph_ay / ph_ax.
// Returns 0, -1 (phases with different tap counts) or -2 (tap count not supported).
static int up_fill_taps(IgemmParams& p, int k, int s, int pad, int big_c) {
    p.num_phases = s * s;
    int per_dim = -1;
    for (int a = 0; a < s; ++a) {
        int n = 0;
        for (int ky = 0; ky < k; ++ky) n += floor_mod(a + pad - ky, s) == 0;
        if (per_dim >= 0 && n != per_dim) return -1;
        per_dim = n;
    }
    const int per_phase = per_dim * per_dim;
    if (per_phase * s * s > 64 || per_phase == 0) return -2;
    for (int ay = 0; ay < s; ++ay)
        for (int ax = 0; ax < s; ++ax) {
            p.ph_ay[ay * s + ax] = ay;
            p.ph_ax[ay * s + ax] = ax;
        }
    p.taps_per_phase = per_phase;
    for (int ay = 0; ay < s; ++ay)
        for (int ax = 0; ax < s; ++ax) {
            const int ph = ay * s + ax;
            int n = 0;
            for (int ky = 0; ky < k; ++ky) {
                if (floor_mod(ay + pad - ky, s) != 0) continue;
                for (int kx = 0; kx < k; ++kx) {
                    if (floor_mod(ax + pad - kx, s) != 0) continue;
                    IgemmTap& t = p.taps[ph * per_phase + n];
                    t.view = 0;
                    t.dy = static_cast<int16_t>((ay + pad - ky) / s);
                    t.dx = static_cast<int16_t>((ax + pad - kx) / s);
                    t.tap_id = static_cast<int16_t>(ky * k + kx);
                    t.brow = (ky * k + kx) * big_c;
                    ++n;
                }
            }
        }
    return 0;
}

// ---------------------------------------------------------------------------------------------- halo tiles
// Experiment switch VG_HALO (off by default).  Regroups the taps of every phase by (view, column
// shift): the taps of a group differ only in their ROW shift, so a pipeline stage loads ONE (16 + hy) x 8 activation tile
// per group and channel chunk and the group's taps read row-shifted windows of it.  A window starts sy * 8 rows = sy
// KB into the tile, i.e. on a 1024-byte swizzle-atom boundary: the operand descriptors stay completely standard.
// (The first variant also shared the tile across COLUMN shifts - windows starting at any 128-byte row, 8-row groups
// a non-1024 stride apart.  It computes correctly but ran 1.45-1.5x SLOWER on every layer (r2b_fused_VG_HALO_64.log):
// the tensor core fetches an 8-row group that straddles two swizzle atoms twice.)
// Needs 128-byte rows (kchunk 64), 16 x 8 output tiles (tb = 1) and an output grid those tiles cover exactly.
// Returns false and leaves `p` alone when the launch does not qualify; on success the caller rebuilds amap[] with
// the (halo_w, halo_h, 1) box and re-derives the stage count.
static bool halo_plan(IgemmParams& p, int grid_w, int grid_h, int batch) {
    // VG_HALO=1: every eligible launch; VG_HALO=<n >= 16>: only launches whose N tile is at most n (e.g. 64)
    const int wanted = switches().halo;
    if (wanted == 0 || (wanted >= 16 && p.n_tile > wanted)) return false;
    if (p.kchunk != 64 || grid_w % 8 != 0 || grid_h % 16 != 0) return false;
    const int tpp = p.taps_per_phase;
    if (tpp < 2 || p.num_phases * tpp > 64) return false;
    IgemmTap ordered[64];
    int16_t org_dy[32], org_dx[32];
    int shift_y[64];
    int gt = -1, groups_total = 0, hy = 0;
    for (int ph = 0; ph < p.num_phases; ++ph) {
        const IgemmTap* src = &p.taps[ph * tpp];
        int n_out = 0, groups_here = 0;
        bool used[64] = {};
        for (int first = 0; first < tpp; ++first) {
            if (used[first]) continue;
            // one group: every not-yet-placed tap of this phase reading the same view at the same column shift
            auto same = [&](int t) { return !used[t] && src[t].view == src[first].view && src[t].dx == src[first].dx; };
            int lo_y = 1 << 20, hi_y = -(1 << 20), n = 0;
            for (int t = first; t < tpp; ++t)
                if (same(t)) {
                    lo_y = std::min<int>(lo_y, src[t].dy); hi_y = std::max<int>(hi_y, src[t].dy);
                    ++n;
                }
            if (gt < 0) gt = n;
            if (n != gt || groups_total >= 32) return false;
            const int16_t gdx = src[first].dx;
            for (int t = first; t < tpp; ++t)
                if (same(t)) {
                    const int o = ph * tpp + n_out++;
                    ordered[o] = src[t];
                    shift_y[o] = src[t].dy - lo_y;
                    used[t] = true;
                }
            org_dy[groups_total] = static_cast<int16_t>(lo_y);
            org_dx[groups_total] = gdx;
            hy = std::max(hy, hi_y - lo_y);
            ++groups_total;
            ++groups_here;
        }
        if (groups_here * gt != tpp) return false;
    }
    // every phase must hold the same number of groups (group index = flat first-tap index / gt)
    if (gt < 2 || gt > 16 || hy > 3 || groups_total * gt != p.num_phases * tpp) return false;
    const int halo_w = 8, halo_h = 16 + hy, row_bytes = p.kchunk * 2;
    const int halo_bytes = halo_w * halo_h * row_bytes;            // (a multiple of 1024)
    const int stage = halo_bytes + gt * p.n_tile * row_bytes;
    if (2 * stage + 4096 + epilogue_extra_smem(p) > 220 * 1024) return false;
    for (int i = 0; i < p.num_phases * tpp; ++i) {
        p.taps[i] = ordered[i];
        p.halo_shift16[i] = static_cast<uint16_t>(shift_y[i] * halo_w * row_bytes / 16);
    }
    for (int g = 0; g < groups_total; ++g) { p.halo_dy[g] = org_dy[g]; p.halo_dx[g] = org_dx[g]; }
    p.halo = 1;
    p.halo_w = halo_w;
    p.halo_h = halo_h;
    p.halo_bytes = halo_bytes;
    p.halo_stage_bytes = halo_bytes;
    p.tps = gt;
    p.b_merged = 0;
    p.tw = 8;
    p.th = 16;
    p.tb = 1;
    p.tiles_w = grid_w / 8;
    p.tiles_h = grid_h / 16;
    p.tiles_b = batch;
    // (two CTAs per SM only if two stages + the epilogue tables fit into half an SM's shared memory)
    const bool half_sm = !owns_sm(p.n_tile) && 2 * stage + 3072 + epilogue_extra_smem(p) <= 112 * 1024;
    p.stages = pick_stages(stage, half_sm ? p.n_tile : 256, 1 << 30, epilogue_extra_smem(p));
    return true;
}

// ---------------------------------------------------------------------------------------------- down
static int pick_ksplit(int tiles, int iters) {
    // few output tiles and a long (tap, channel-chunk) loop: spread the reduction over ~one wave of CTAs
    if (tiles >= 74 || iters < 32) return 1;
    return std::max(1, std::min(iters / 8, 148 / tiles));
}

static int down_umma(const VgConvGeom* g, const void* big, const void* wd, const float* bias, void* small,
                     int out_f32, void* ws, size_t ws_bytes, cudaStream_t stream, const VgEpilogue* ep = nullptr) {
    if (!aligned16(big) || !aligned16(wd) || !aligned16(small)) return fail(VG_ERR_ALIGN, "down: 16-byte alignment");
    IgemmParams p;
    std::memset(&p, 0, sizeof(p));
    fuse_fill(p, g, ep);
    const int k = g->kernel, s = g->stride, pad = g->pad;
    p.kchunk = pick_kchunk(g->big_c);
    p.c_chunks = g->big_c / p.kchunk;
    down_tiling(g, &p.tw, &p.th, &p.tb, &p.n_tile);
    p.n_tiles = g->small_c / p.n_tile;
    p.tiles_w = ceil_div(g->small_w, p.tw);
    p.tiles_h = ceil_div(g->small_h, p.th);
    p.tiles_b = ceil_div(g->batch, p.tb);
    p.num_phases = 1;
    p.taps_per_phase = k * k;
    const int swz = p.kchunk * 2;
    const int nviews = s * s;
    for (int v = 0; v < 4; ++v) {
        const int vv = v < nviews ? v : 0;
        const int rc = make_view(&p.amap[v], big, g->batch, g->big_h, g->big_w, g->big_c, s, vv / s, vv % s, p.kchunk,
                                 p.tw, p.th, p.tb, swz);
        if (rc != 0) return fail(VG_ERR_CUDA, "down: cuTensorMapEncodeTiled(A view %d) failed (%d)", v, rc);
    }
    p.tps = pick_tps(p.kchunk, p.c_chunks, p.taps_per_phase);
    // taps are enumerated in packed order, so with a single N tile the tps slabs of a stage are adjacent rows
    p.b_merged = p.tps > 1 && p.n_tiles == 1 && p.tps * p.n_tile <= 256;
    {
        const uint64_t dims[2] = {(uint64_t)g->big_c, (uint64_t)k * k * g->small_c};
        const uint64_t strides[2] = {1, (uint64_t)g->big_c};
        const uint32_t box[2] = {(uint32_t)p.kchunk, (uint32_t)(p.b_merged ? p.tps * p.n_tile : p.n_tile)};
        const int rc = make_tmap_bf16(&p.bmap, wd, 2, dims, strides, box, swz);
        if (rc != 0) return fail(VG_ERR_CUDA, "down: cuTensorMapEncodeTiled(B) failed (%d)", rc);
    }
    down_fill_taps(p, k, s, pad, g->small_c);
    p.tps = pick_tps(p.kchunk, p.c_chunks, p.taps_per_phase);
    p.out = small;
    p.out_fp32 = out_f32;
    p.out_B = g->batch;
    p.out_H = g->small_h;
    p.out_W = g->small_w;
    p.out_C = g->small_c;
    p.osy = p.osx = 1;
    p.bias = bias;
    const size_t acc_bytes = static_cast<size_t>(g->batch) * g->small_h * g->small_w * g->small_c * sizeof(float);
    const int ks = pick_ksplit(p.tiles_w * p.tiles_h * p.tiles_b * p.n_tiles, p.taps_per_phase * p.c_chunks);
    if (ks > 1 && p.fuse_mode == 0 && ws != nullptr && ws_bytes >= acc_bytes && (reinterpret_cast<uintptr_t>(ws) & 15) == 0) {
        p.ksplit = ks;
        p.splitk_acc = static_cast<float*>(ws);
    }
    plan_epilogue_flags(p, small, g->small_c, out_f32);
    p.stages = pick_stages(p.tps * (128 + p.n_tile) * p.kchunk * 2, p.n_tile, p.taps_per_phase / p.tps * p.c_chunks,
                           epilogue_extra_smem(p));
    if (p.ksplit <= 1 && halo_plan(p, g->small_w, g->small_h, g->batch)) {
        for (int v = 0; v < 4; ++v) {
            const int vv = v < nviews ? v : 0;
            const int rc = make_view(&p.amap[v], big, g->batch, g->big_h, g->big_w, g->big_c, s, vv / s, vv % s, p.kchunk,
                                     p.halo_w, p.halo_h, 1, swz);
            if (rc != 0) return fail(VG_ERR_CUDA, "down: cuTensorMapEncodeTiled(A halo view %d) failed (%d)", v, rc);
        }
        const uint64_t dims[2] = {(uint64_t)g->big_c, (uint64_t)k * k * g->small_c};
        const uint64_t strides[2] = {1, (uint64_t)g->big_c};
        const uint32_t box[2] = {(uint32_t)p.kchunk, (uint32_t)p.n_tile};
        const int rc = make_tmap_bf16(&p.bmap, wd, 2, dims, strides, box, swz);
        if (rc != 0) return fail(VG_ERR_CUDA, "down: cuTensorMapEncodeTiled(B) failed (%d)", rc);
    }
    if (plan_epilogue_maps(p, small, g->batch, g->small_h, g->small_w, g->small_c) != 0)
        return fail(VG_ERR_CUDA, "down: cuTensorMapEncodeTiled(output) failed");
    const int rc = launch_igemm(p, stream);
    if (rc != 0) return cuda_fail(static_cast<cudaError_t>(rc), "igemm_fprop_kernel<down>");
    note_launch(p.ksplit > 1 ? 2 : 1);
    return VG_OK;
}

// ---------------------------------------------------------------------------------------------- up
static int up_umma(const VgConvGeom* g, const void* small, const void* wu, void* big, cudaStream_t stream,
                   const VgEpilogue* ep = nullptr) {
    if (!aligned16(big) || !aligned16(wu) || !aligned16(small)) return fail(VG_ERR_ALIGN, "up: 16-byte alignment");
    IgemmParams p;
    std::memset(&p, 0, sizeof(p));
    fuse_fill(p, g, ep);
    const int k = g->kernel, s = g->stride, pad = g->pad;
    const bool dense = g->small_h == 1 && g->small_w == 1 && s == 1 && pad == 0;
    p.kchunk = pick_kchunk(g->small_c);
    p.c_chunks = g->small_c / p.kchunk;
    const int swz = p.kchunk * 2;
    int n_total = 0;
    up_tiling(g, &p.tw, &p.th, &p.tb, &p.n_tile, &n_total);
    p.n_tiles = n_total / p.n_tile;
    const int grid_h = dense ? 1 : ceil_div(g->big_h, s), grid_w = dense ? 1 : ceil_div(g->big_w, s);
    p.tiles_w = ceil_div(grid_w, p.tw);
    p.tiles_h = ceil_div(grid_h, p.th);
    p.tiles_b = ceil_div(g->batch, p.tb);
    for (int v = 0; v < 4; ++v) {
        const int rc = make_view(&p.amap[v], small, g->batch, g->small_h, g->small_w, g->small_c, 1, 0, 0, p.kchunk,
                                 p.tw, p.th, p.tb, swz);
        if (rc != 0) return fail(VG_ERR_CUDA, "up: cuTensorMapEncodeTiled(A) failed (%d)", rc);
    }
    {
        const uint64_t dims[2] = {(uint64_t)g->small_c, (uint64_t)k * k * g->big_c};
        const uint64_t strides[2] = {1, (uint64_t)g->small_c};
        const uint32_t box[2] = {(uint32_t)p.kchunk, (uint32_t)p.n_tile};
        const int rc = make_tmap_bf16(&p.bmap, wu, 2, dims, strides, box, swz);
        if (rc != 0) return fail(VG_ERR_CUDA, "up: cuTensorMapEncodeTiled(B) failed (%d)", rc);
    }
    p.out = big;
    p.out_fp32 = 0;
    p.out_B = g->batch;
    if (dense) {
        // big[b, (ky,kx), bc] = sum_sc small[b, sc] * w[sc][bc][ky][kx]: one GEMM with N = k*k*big_c
        p.num_phases = 1;
        p.taps_per_phase = 1;
        p.taps[0] = IgemmTap{0, 0, 0, 0, 0};
        p.out_H = 1;
        p.out_W = 1;
        p.out_C = n_total;
        p.osy = p.osx = 1;
    } else {
        const int trc = up_fill_taps(p, k, s, pad, g->big_c);
        if (trc == -1) return fail(VG_ERR_SHAPE, "up: phases with different tap counts (k%d s%d)", k, s);
        if (trc == -2) return fail(VG_ERR_SHAPE, "up: unsupported tap count");
        p.out_H = g->big_h;
        p.out_W = g->big_w;
        p.out_C = g->big_c;
        p.osy = p.osx = s;
    }
    p.tps = pick_tps(p.kchunk, p.c_chunks, p.taps_per_phase);
    plan_epilogue_flags(p, big, p.out_C, 0);
    p.stages = pick_stages(p.tps * (128 + p.n_tile) * p.kchunk * 2, p.n_tile, p.taps_per_phase / p.tps * p.c_chunks,
                           epilogue_extra_smem(p));
    if (!dense && halo_plan(p, grid_w, grid_h, g->batch)) {
        for (int v = 0; v < 4; ++v) {
            const int rc = make_view(&p.amap[v], small, g->batch, g->small_h, g->small_w, g->small_c, 1, 0, 0, p.kchunk,
                                     p.halo_w, p.halo_h, 1, swz);
            if (rc != 0) return fail(VG_ERR_CUDA, "up: cuTensorMapEncodeTiled(A halo) failed (%d)", rc);
        }
    }
    if (plan_epilogue_maps(p, big, g->batch, p.out_H, p.out_W, p.out_C) != 0)
        return fail(VG_ERR_CUDA, "up: cuTensorMapEncodeTiled(output) failed");
    const int rc = launch_igemm(p, stream);
    if (rc != 0) return cuda_fail(static_cast<cudaError_t>(rc), "igemm_fprop_kernel<up>");
    note_launch();
    return VG_OK;
}

// ---------------------------------------------------------------------------------------------- wgrad
static int wgrad_umma(const VgConvGeom* g, const void* small, const void* big, float* dw, void* ws, size_t ws_bytes,
                      cudaStream_t stream, size_t* ws_needed = nullptr, int overwrite = 0) {
    if (ws_needed == nullptr && (!aligned16(big) || !aligned16(small)))
        return fail(VG_ERR_ALIGN, "wgrad: 16-byte alignment");
    WgradParams p;
    std::memset(&p, 0, sizeof(p));
    const int k = g->kernel, s = g->stride, pad = g->pad;
    // taps per UMMA (side by side along N, N <= 256) and pixels reduced per pipeline stage (Q stage ~32 KB)
    const int nt = wgrad_n_tile(g->big_c);
    const int tpc = std::min(std::min(k * k, 512 / nt), 16);
    const int merge = nt <= 64 ? std::max(1, std::min(tpc, 256 / nt)) : 1;   // wide tiles keep one tap per UMMA
    const int kpix = switches().wgrad_kpix;
    pick_box(kpix, g->small_w, g->small_h, &p.tw, &p.th, &p.tb);
    p.tiles_w = ceil_div(g->small_w, p.tw);
    p.tiles_h = ceil_div(g->small_h, p.th);
    p.tiles_b = ceil_div(g->batch, p.tb);
    p.p_atom_c = g->small_c % 64 == 0 ? 64 : (g->small_c % 32 == 0 ? 32 : 16);
    p.m_atoms = std::min(128, g->small_c) / p.p_atom_c;
    p.m_tiles = ceil_div(g->small_c, p.m_atoms * p.p_atom_c);
    p.n_tile = wgrad_n_tile(g->big_c);
    p.q_atom_c = std::min(64, p.n_tile);
    p.n_tiles = g->big_c / p.n_tile;
    p.num_taps = k * k;
    p.taps_per_cta = std::min(p.num_taps, 512 / p.n_tile);
    if (p.num_taps % 4 == 0) p.taps_per_cta = std::max(4, p.taps_per_cta / 4 * 4);
    p.taps_per_cta = std::min(p.taps_per_cta, 16);
    p.merge = p.n_tile <= 64 ? std::max(1, std::min(p.taps_per_cta, 256 / p.n_tile)) : 1;
    if (ws_needed == nullptr) {
        const int rc = make_view(&p.pmap, small, g->batch, g->small_h, g->small_w, g->small_c, 1, 0, 0, p.p_atom_c, p.tw,
                                 p.th, p.tb, p.p_atom_c * 2);
        if (rc != 0) return fail(VG_ERR_CUDA, "wgrad: cuTensorMapEncodeTiled(P) failed (%d)", rc);
    }
    const int nviews = s * s;
    for (int v = 0; v < 4 && ws_needed == nullptr; ++v) {
        const int vv = v < nviews ? v : 0;
        const int rc = make_view(&p.qmap[v], big, g->batch, g->big_h, g->big_w, g->big_c, s, vv / s, vv % s, p.q_atom_c,
                                 p.tw, p.th, p.tb, p.q_atom_c * 2);
        if (rc != 0) return fail(VG_ERR_CUDA, "wgrad: cuTensorMapEncodeTiled(Q view %d) failed (%d)", v, rc);
    }
    for (int ky = 0; ky < k; ++ky)
        for (int kx = 0; kx < k; ++kx) {
            IgemmTap& t = p.taps[ky * k + kx];
            const int vy = floor_mod(ky - pad, s), vx = floor_mod(kx - pad, s);
            t.view = static_cast<int16_t>(vy * s + vx);
            t.dy = static_cast<int16_t>((ky - pad - vy) / s);
            t.dx = static_cast<int16_t>((kx - pad - vx) / s);
            t.tap_id = static_cast<int16_t>(ky * k + kx);
            t.brow = 0;
        }
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    const int tap_groups = ceil_div(p.num_taps, p.taps_per_cta);
    const int base_ctas = p.m_tiles * p.n_tiles * tap_groups;
    // one CTA per SM (the accumulators fill TMEM): aim at a single full wave; split the pixel range only when the
    // (channel tile, tap group) grid alone leaves most SMs idle
    p.splits = base_ctas >= 100 ? 1 : std::max(1, std::min(total_tiles, 148 / base_ctas));
    const int a_stage = kpix * 256, b_stage = p.merge * p.n_tile * kpix * 2;
    p.stages_a = kpix >= 128 ? 2 : 3;
    if (switches().wgrad_sa > 0) p.stages_a = std::max(2, switches().wgrad_sa);      // experiment switch
    p.stages_b = std::max(2, std::min(20, (200 * 1024 - p.stages_a * a_stage) / b_stage));
    p.dw = dw;
    p.debug_flags = switches().debug_wgrad;
    p.accumulate = overwrite ? 0 : 1;      // (overwrite: 1 = unknown content, 2 = the caller says dw is zero)
    const int bcv = bc_valid(g);
    p.s_m = static_cast<long long>(bcv) * k * k;
    p.s_n = k * k;
    p.s_tap = 1;
    p.m_valid = g->small_c;
    p.n_valid = bcv;
    // dw[m][n][tap..tap+3] contiguous and 16-byte aligned when k*k is a multiple of 4
    p.vec4_taps = (k * k) % 4 == 0 && p.taps_per_cta % 4 == 0 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0 &&
                  (p.s_m % 4) == 0;
    {
        // CTA pairs sharing the Q operand (experiment switch VG_WGRAD2=1)
        const bool want_pair = switches().wgrad_pair != 0;
        const int n_atoms = p.n_tile / p.q_atom_c;
        p.pair = want_pair && p.m_tiles % 2 == 0 && p.m_atoms * p.p_atom_c == 128 && g->small_c % 128 == 0 &&
                 (p.merge * n_atoms) % 2 == 0 && p.taps_per_cta % p.merge == 0 && p.num_taps % p.taps_per_cta == 0;
        if (p.pair) {
            // each CTA of a pair holds half of every Q stage: the same shared memory carries a ring twice as deep
            const int a_stage_b = p.tw * p.th * p.tb * 256, b_half = p.merge * p.n_tile * p.tw * p.th * p.tb;
            p.stages_b = std::max(2, std::min(20, (200 * 1024 - p.stages_a * a_stage_b) / b_half));
        }
    }
    // Split layers (few weight tiles, a long pixel reduction): every split adds its tile straight into dw with 16-byte
    // vector reductions (red.global.add.v4.f32; dw is small and L2-resident) - no partial tiles, no second kernel.
    p.atomic_split = p.splits > 1 && p.vec4_taps && switches().wgrad_atomic != 0;
    if (p.atomic_split) {
        if (ws_needed != nullptr) { *ws_needed = 0; return VG_OK; }
    } else if (p.splits > 1) {
        const size_t need = wgrad_partial_bytes(p);
        if (ws_needed != nullptr) { *ws_needed = need; return VG_OK; }
        if (ws != nullptr && ws_bytes >= need && (reinterpret_cast<uintptr_t>(ws) & 15) == 0) p.partial = static_cast<float*>(ws);
        else p.splits = 1;   // no scratch: correct but with idle SMs
    } else if (ws_needed != nullptr) {
        *ws_needed = 0;
        return VG_OK;
    }
    if (p.splits > 1 && overwrite) {       // every split path accumulates: start from zero
        if (overwrite == 1) {
            const cudaError_t e =
                cudaMemsetAsync(dw, 0, static_cast<size_t>(g->small_c) * bcv * k * k * sizeof(float), stream);
            if (e != cudaSuccess) return cuda_fail(e, "wgrad: memset");
        }
        p.accumulate = 1;
    }
    const int rc = launch_wgrad(p, stream);
    if (rc != 0) return cuda_fail(static_cast<cudaError_t>(rc), "igemm_wgrad_kernel");
    note_launch((p.splits > 1 && !p.atomic_split) ? 2 : 1);
    return VG_OK;
}

// ---------------------------------------------------------------------------------------------- packing
// fp32 master w[small_c][big_c_valid][kk]  ->  bf16 wd[tap][small_c][big_c] and wu[tap][big_c][small_c].
// One block moves a 32(small_c) x 32(big_c) x <=16(tap) tile through shared memory: the master is read in contiguous
// 2 KB runs and both copies are written in 64-byte runs (the first version gathered with a 64-byte stride and
// scattered 2-byte stores: 8 B/parameter of useful traffic moved at ~5 % of HBM speed).
struct PackBatch {
    VgPackItem item[16];
    int n;
};

// (4 taps per tile: the discriminator's largest layer then has 512 tiles instead of 128 - with 16 taps per tile one
// block per SM moved the whole 17 MB and the launch sat on the critical path for 16 us after every discriminator Adam)
constexpr int kPackT = 32, kPackK = 4, kPackRow = kPackT + 2, kPackPlane = kPackT * kPackRow + 2;

__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const PackBatch b) {
    pdl_enter();
    __shared__ __nv_bfloat16 tile[kPackK * kPackPlane];
    const VgPackItem it = b.item[blockIdx.y];
    const int sc = it.small_c, bc = it.big_c, bcv = it.big_c_valid > 0 ? it.big_c_valid : it.big_c, kk = it.kk;
    __nv_bfloat16* wd = static_cast<__nv_bfloat16*>(it.wd);
    __nv_bfloat16* wu = static_cast<__nv_bfloat16*>(it.wu);
    if ((sc | bc) & 1) {
        // odd extents (e.g. the discriminator's single-output head): element-wise, destination-ordered
        const long long n = static_cast<long long>(sc) * bc * kk;
        for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
             i += static_cast<long long>(gridDim.x) * blockDim.x) {
            const int bi = static_cast<int>(i % bc);
            const int s = static_cast<int>((i / bc) % sc);
            const int tap = static_cast<int>(i / (static_cast<long long>(bc) * sc));
            const float v = bi < bcv ? it.w[(static_cast<long long>(s) * bcv + bi) * kk + tap] : 0.f;
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            if (wd != nullptr) wd[i] = h;
            if (wu != nullptr) wu[(static_cast<long long>(tap) * bc + bi) * sc + s] = h;
        }
        return;
    }
    const int tiles_b = (bc + kPackT - 1) / kPackT, tiles_s = (sc + kPackT - 1) / kPackT;
    const int kchunks = (kk + kPackK - 1) / kPackK;
    const long long work = static_cast<long long>(tiles_b) * tiles_s * kchunks;
    for (long long t = blockIdx.x; t < work; t += gridDim.x) {
        const int kc = static_cast<int>(t % kchunks);
        const int tb = static_cast<int>((t / kchunks) % tiles_b);
        const int ts = static_cast<int>(t / (static_cast<long long>(kchunks) * tiles_b));
        const int s0 = ts * kPackT, b0 = tb * kPackT, k0 = kc * kPackK, nk = min(kPackK, kk - k0);
        if (kk == 16 && (reinterpret_cast<uintptr_t>(it.w) & 15) == 0) {
            // 4x4 kernels (every stride-2 layer): one 16-byte load = the tile's four taps, no integer division
            static_assert(kPackK == 4, "the 16-byte path takes one float4 (four taps) per (small_c, big_c) pair");
            for (int e = threadIdx.x; e < kPackT * kPackT; e += blockDim.x) {
                const int s_l = e >> 5, bi_l = e & 31;
                const int s = s0 + s_l, bi = b0 + bi_l;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (s < sc && bi < bcv)
                    v = __ldg(reinterpret_cast<const float4*>(it.w + (static_cast<long long>(s) * bcv + bi) * 16) + kc);
                __nv_bfloat16* t4 = tile + s_l * kPackRow + bi_l;
                t4[0] = __float2bfloat16_rn(v.x);
                t4[kPackPlane] = __float2bfloat16_rn(v.y);
                t4[2 * kPackPlane] = __float2bfloat16_rn(v.z);
                t4[3 * kPackPlane] = __float2bfloat16_rn(v.w);
            }
        } else {
            const int span = kPackT * nk;                   // (big_c, tap) elements of one small_c row of the tile
            for (int e = threadIdx.x; e < kPackT * span; e += blockDim.x) {
                const int s_l = e / span, rem = e - s_l * span;
                const int bi_l = rem / nk, tp = rem - bi_l * nk;
                const int s = s0 + s_l, bi = b0 + bi_l;
                float v = 0.f;
                if (s < sc && bi < bcv) v = __ldg(it.w + (static_cast<long long>(s) * bcv + bi) * kk + k0 + tp);
                tile[tp * kPackPlane + s_l * kPackRow + bi_l] = __float2bfloat16_rn(v);
            }
        }
        __syncthreads();
        constexpr int kHalf = kPackT / 2;
        if (wd != nullptr) {
            for (int e = threadIdx.x; e < nk * kPackT * kHalf; e += blockDim.x) {
                const int bp = e & (kHalf - 1), s_l = (e >> 4) & (kPackT - 1), tp = e >> 9;
                const int s = s0 + s_l, bi = b0 + 2 * bp;
                if (s < sc && bi < bc)
                    *reinterpret_cast<__nv_bfloat162*>(wd + (static_cast<long long>(k0 + tp) * sc + s) * bc + bi) =
                        *reinterpret_cast<const __nv_bfloat162*>(&tile[tp * kPackPlane + s_l * kPackRow + 2 * bp]);
            }
        }
        if (wu != nullptr) {
            for (int e = threadIdx.x; e < nk * kPackT * kHalf; e += blockDim.x) {
                const int sp = e & (kHalf - 1), bi_l = (e >> 4) & (kPackT - 1), tp = e >> 9;
                const int s = s0 + 2 * sp, bi = b0 + bi_l;
                if (s < sc && bi < bc) {
                    __nv_bfloat162 h;
                    h.x = tile[tp * kPackPlane + (2 * sp) * kPackRow + bi_l];
                    h.y = tile[tp * kPackPlane + (2 * sp + 1) * kPackRow + bi_l];
                    *reinterpret_cast<__nv_bfloat162*>(wu + (static_cast<long long>(k0 + tp) * bc + bi) * sc + s) = h;
                }
            }
        }
        __syncthreads();
    }
}

static long long pack_work(const VgPackItem& it) {
    if ((it.small_c | it.big_c) & 1) return (static_cast<long long>(it.small_c) * it.big_c * it.kk + 255) / 256;
    return static_cast<long long>((it.big_c + kPackT - 1) / kPackT) * ((it.small_c + kPackT - 1) / kPackT) *
           ((it.kk + kPackK - 1) / kPackK);
}

}  // namespace vg

using namespace vg;

extern "C" int vg_pack_weights_multi(const VgPackItem* items, int n_items, void* stream) {
    if (items == nullptr || n_items <= 0) return fail(VG_ERR_ARG, "pack_multi: no items");
    int rc = device_check();
    if (rc != VG_OK) return rc;
    for (int base = 0; base < n_items; base += 16) {
        PackBatch b;
        b.n = std::min(16, n_items - base);
        long long biggest = 0;
        for (int i = 0; i < b.n; ++i) {
            b.item[i] = items[base + i];
            if (b.item[i].w == nullptr) return fail(VG_ERR_ARG, "pack_multi: null master weight");
            biggest = std::max(biggest, pack_work(b.item[i]));
        }
        const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>(biggest, 148 * 6)));
        launch_k(pack_weights_multi_kernel, dim3(blocks, b.n), dim3(256), 0, as_stream(stream), b);
        VG_LAUNCHED();
    }
    return VG_OK;
}

extern "C" int vg_pack_weights_bf16(const VgConvGeom* g, const float* w, void* wd, void* wu, void* stream) {
    if (g == nullptr || w == nullptr) return fail(VG_ERR_ARG, "pack: null argument");
    int rc = device_check();
    if (rc != VG_OK) return rc;
    PackBatch b;
    b.n = 1;
    b.item[0].w = w;
    b.item[0].wd = wd;
    b.item[0].wu = wu;
    b.item[0].small_c = g->small_c;
    b.item[0].big_c = g->big_c;
    b.item[0].big_c_valid = bc_valid(g) != g->big_c ? bc_valid(g) : 0;
    b.item[0].kk = g->kernel * g->kernel;
    const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>(pack_work(b.item[0]), 148 * 6)));
    launch_k(pack_weights_multi_kernel, dim3(blocks, 1), dim3(256), 0, as_stream(stream), b);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" size_t vg_conv_down_workspace_bytes(const VgConvGeom* g) {
    if (g == nullptr) return 0;
    return static_cast<size_t>(g->batch) * g->small_h * g->small_w * g->small_c * sizeof(float);
}

extern "C" int vg_conv_down(const VgConvGeom* g, VgDType dtype, const void* big, const void* w, const float* bias,
                            void* small, int out_f32, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_geom(g);
    if (rc != VG_OK) return rc;
    if (big == nullptr || w == nullptr || small == nullptr) return fail(VG_ERR_ARG, "down: null pointer");
    rc = device_check();
    if (rc != VG_OK) return rc;
    if (is_gemv(g)) return gemv_down(g, dtype, big, w, bias, small, out_f32, as_stream(stream));
    if (dtype == VG_BF16 && umma_down_ok(g))
        return down_umma(g, big, w, bias, small, out_f32, ws, ws_bytes, as_stream(stream));
    return simt_conv_down(g, dtype, big, w, bias, small, out_f32, as_stream(stream));
}

extern "C" int vg_conv_up(const VgConvGeom* g, VgDType dtype, const void* small, const void* w, void* big,
                          void* stream) {
    int rc = check_geom(g);
    if (rc != VG_OK) return rc;
    if (big == nullptr || w == nullptr || small == nullptr) return fail(VG_ERR_ARG, "up: null pointer");
    rc = device_check();
    if (rc != VG_OK) return rc;
    if (is_gemv(g)) return gemv_up(g, dtype, small, w, big, as_stream(stream));
    if (dtype == VG_BF16 && umma_up_ok(g)) return up_umma(g, small, w, big, as_stream(stream));
    return simt_conv_up(g, dtype, small, w, big, as_stream(stream));
}

extern "C" int vg_conv_epilogue_supported(const VgConvGeom* g, VgDType dtype, int up, const VgEpilogue* ep) {
    if (check_geom(g) != VG_OK || dtype != VG_BF16 || ep == nullptr) return 0;
    return fuse_check(g, up != 0, ep) == VG_OK ? 1 : 0;
}

extern "C" int vg_conv_down_ex(const VgConvGeom* g, VgDType dtype, const void* big, const void* w, const float* bias,
                               void* small, const VgEpilogue* ep, void* stream) {
    int rc = check_geom(g);
    if (rc != VG_OK) return rc;
    if (big == nullptr || w == nullptr || small == nullptr) return fail(VG_ERR_ARG, "down: null pointer");
    rc = device_check();
    if (rc != VG_OK) return rc;
    if (dtype != VG_BF16) return fail(VG_ERR_SHAPE, "fused epilogues exist on the bf16 tensor-core path only");
    rc = fuse_check(g, false, ep);
    if (rc != VG_OK) return rc;
    return down_umma(g, big, w, bias, small, 0, nullptr, 0, as_stream(stream), ep);
}

extern "C" int vg_conv_up_ex(const VgConvGeom* g, VgDType dtype, const void* small, const void* w, void* big,
                             const VgEpilogue* ep, void* stream) {
    int rc = check_geom(g);
    if (rc != VG_OK) return rc;
    if (big == nullptr || w == nullptr || small == nullptr) return fail(VG_ERR_ARG, "up: null pointer");
    rc = device_check();
    if (rc != VG_OK) return rc;
    if (dtype != VG_BF16) return fail(VG_ERR_SHAPE, "fused epilogues exist on the bf16 tensor-core path only");
    rc = fuse_check(g, true, ep);
    if (rc != VG_OK) return rc;
    if (is_gemv(g) && ep != nullptr && ep->mode != VG_EPI_NONE) {
        if (!aligned16(w) || !aligned16(big)) return fail(VG_ERR_ALIGN, "up: 16-byte alignment");
        return gemv_up_fused(g, small, w, big, ep, as_stream(stream));
    }
    if (is_gemv(g)) return gemv_up(g, dtype, small, w, big, as_stream(stream));
    return up_umma(g, small, w, big, as_stream(stream), ep);
}

extern "C" size_t vg_conv_wgrad_workspace_bytes(const VgConvGeom* g, VgDType dtype) {
    if (g == nullptr || check_geom(g) != VG_OK || dtype != VG_BF16 || is_gemv(g) || !umma_wgrad_ok(g)) return 0;
    size_t need = 0;
    float dummy = 0.f;
    wgrad_umma(g, &dummy, &dummy, &dummy, nullptr, 0, nullptr, &need);
    return need;
}

static int wgrad_dispatch(const VgConvGeom* g, VgDType dtype, const void* small, const void* big, float* dw, void* ws,
                          size_t ws_bytes, int overwrite, void* stream) {
    int rc = check_geom(g);
    if (rc != VG_OK) return rc;
    if (big == nullptr || dw == nullptr || small == nullptr) return fail(VG_ERR_ARG, "wgrad: null pointer");
    rc = device_check();
    if (rc != VG_OK) return rc;
    if (dtype == VG_BF16 && !is_gemv(g) && umma_wgrad_ok(g))
        return wgrad_umma(g, small, big, dw, ws, ws_bytes, as_stream(stream), nullptr, overwrite);
    if (overwrite == 1) {     // the other paths accumulate: give them a zeroed destination
        const size_t n = static_cast<size_t>(g->small_c) * bc_valid(g) * g->kernel * g->kernel;
        const cudaError_t e = cudaMemsetAsync(dw, 0, n * sizeof(float), as_stream(stream));
        if (e != cudaSuccess) return cuda_fail(e, "wgrad: memset");
    }
    if (is_gemv(g)) return gemv_wgrad(g, dtype, small, big, dw, as_stream(stream));
    return simt_conv_wgrad(g, dtype, small, big, dw, as_stream(stream));
}

extern "C" int vg_conv_wgrad(const VgConvGeom* g, VgDType dtype, const void* small, const void* big, float* dw,
                             void* ws, size_t ws_bytes, void* stream) {
    return wgrad_dispatch(g, dtype, small, big, dw, ws, ws_bytes, 0, stream);
}

extern "C" int vg_conv_wgrad_ex(const VgConvGeom* g, VgDType dtype, const void* small, const void* big, float* dw,
                                void* ws, size_t ws_bytes, int flags, void* stream) {
    if (flags & ~(VG_WGRAD_OVERWRITE | VG_WGRAD_DST_ZERO)) return fail(VG_ERR_ARG, "wgrad: unknown flags 0x%x", flags);
    return wgrad_dispatch(g, dtype, small, big, dw, ws, ws_bytes,
                          (flags & VG_WGRAD_OVERWRITE) ? 1 : ((flags & VG_WGRAD_DST_ZERO) ? 2 : 0), stream);
}
