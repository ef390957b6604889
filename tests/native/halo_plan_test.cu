// CPU-only check (test infrastructure) of the halo-tile plan in csrc/conv_api.cu: for several convolution geometries it
// builds the tap list exactly as the launchers do, lets halo_plan() regroup it, and then evaluates the contraction
// twice on the host -
//   (a) tap by tap, the way the per-tap kernel reads shifted 128-pixel boxes of the parity views, and
//   (b) the way the HALO kernel does: one (16 + hy) x (8 + hx) tile per group, window row m of tap t =
//       halo row  shift(t) + (m / 8) * halo_w + m % 8  (igemm_umma.cu, tests/native/umma_halo.cu)
// - and requires identical results (small integer data, exact in float).  No CUDA call is made.
// Build: see tests/test_halo_plan_cpu.py (nvcc, linked against the library's other objects).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../vae-gan-based-model-for-image-generation-and-denoising_b200/csrc/conv_api.cu"

using namespace vg;

struct Geo { const char* name; bool up; int B, big_h, big_w, k, s, pad; };

// parity view v of a [B][H][W] plane (channel handled by the caller); zero outside the view, like TMA's fill
struct Views {
    int H, W, s;
    const std::vector<float>* data;     // [B][H][W][C]
    int C;
    float at(int view, int b, int y, int x, int c) const {
        const int vy = view / s, vx = view % s;
        const int Hv = (H - vy + s - 1) / s, Wv = (W - vx + s - 1) / s;
        if (y < 0 || x < 0 || y >= Hv || x >= Wv) return 0.f;
        return (*data)[((static_cast<size_t>(b) * H + (vy + s * y)) * W + (vx + s * x)) * C + c];
    }
};

static int run(const Geo& g) {
    const int C = 2, N = 3;              // the plan does not depend on the channel counts; keep the host loops small
    const int small_h = (g.big_h + 2 * g.pad - g.k) / g.s + 1, small_w = (g.big_w + 2 * g.pad - g.k) / g.s + 1;
    IgemmParams p;
    std::memset(&p, 0, sizeof(p));
    p.kchunk = 64;                       // what the eligibility test looks at
    p.c_chunks = 1;
    p.n_tile = 64;
    int grid_h, grid_w, in_h, in_w, view_s;
    if (!g.up) {
        p.num_phases = 1;
        p.taps_per_phase = g.k * g.k;
        down_fill_taps(p, g.k, g.s, g.pad, N);
        grid_h = small_h; grid_w = small_w; in_h = g.big_h; in_w = g.big_w; view_s = g.s;
    } else {
        if (up_fill_taps(p, g.k, g.s, g.pad, N) != 0) { printf("%s: tap list rejected\n", g.name); return 1; }
        grid_h = (g.big_h + g.s - 1) / g.s; grid_w = (g.big_w + g.s - 1) / g.s; in_h = small_h; in_w = small_w; view_s = 1;
    }
    const IgemmParams before = p;
    if (!halo_plan(p, grid_w, grid_h, g.B)) { printf("%s: not eligible (unexpected)\n", g.name); return 1; }
    const int tpp = p.taps_per_phase, gt = p.tps;
    if (p.tw != 8 || p.th != 16 || p.tb != 1 || p.tiles_w * 8 != grid_w || p.tiles_h * 16 != grid_h || p.tiles_b != g.B ||
        p.halo_bytes != p.halo_w * p.halo_h * 128 || p.halo_stage_bytes % 1024 != 0 || p.halo_stage_bytes < p.halo_bytes ||
        tpp % gt != 0 || p.stages < 2) {
        printf("%s: inconsistent plan\n", g.name);
        return 1;
    }
    // data: small integers
    std::vector<float> x(static_cast<size_t>(g.B) * in_h * in_w * C), w(static_cast<size_t>(g.k) * g.k * N * C);
    unsigned seed = 12345u;
    auto rnd = [&] { seed = seed * 1664525u + 1013904223u; return static_cast<float>(static_cast<int>((seed >> 24) % 7) - 3); };
    for (float& v : x) v = rnd();
    for (float& v : w) v = rnd();
    const Views V{in_h, in_w, view_s, &x, C};
    long long mismatches = 0, checked = 0;
    for (int ph = 0; ph < p.num_phases; ++ph)
        for (int b = 0; b < g.B; ++b)
            for (int ti = 0; ti < p.tiles_h; ++ti)
                for (int tj = 0; tj < p.tiles_w; ++tj) {
                    const int i0 = ti * 16, j0 = tj * 8;
                    for (int n = 0; n < N; ++n) {
                        float ref[128] = {}, halo[128] = {};
                        // (a) the per-tap formulation with the ORIGINAL tap list
                        for (int t = 0; t < tpp; ++t) {
                            const IgemmTap& tap = before.taps[ph * tpp + t];
                            for (int m = 0; m < 128; ++m)
                                for (int c = 0; c < C; ++c)
                                    ref[m] += V.at(tap.view, b, i0 + m / 8 + tap.dy, j0 + m % 8 + tap.dx, c) *
                                              w[(static_cast<size_t>(tap.brow) + n) * C + c];
                        }
                        // (b) the halo formulation with the regrouped list
                        for (int grp = 0; grp < tpp / gt; ++grp) {
                            const int first = grp * gt, gi = (ph * tpp + first) / gt;
                            const int view = p.taps[ph * tpp + first].view;
                            std::vector<float> tile(static_cast<size_t>(p.halo_w) * p.halo_h * C);
                            for (int hy = 0; hy < p.halo_h; ++hy)
                                for (int hx = 0; hx < p.halo_w; ++hx)
                                    for (int c = 0; c < C; ++c)
                                        tile[(static_cast<size_t>(hy) * p.halo_w + hx) * C + c] =
                                            V.at(view, b, i0 + p.halo_dy[gi] + hy, j0 + p.halo_dx[gi] + hx, c);
                            for (int t = 0; t < gt; ++t) {
                                const IgemmTap& tap = p.taps[ph * tpp + first + t];
                                if (tap.view != view) { printf("%s: mixed views in a group\n", g.name); return 1; }
                                const int shift_rows = p.halo_shift16[ph * tpp + first + t] / 8;     // 128-byte rows
                                for (int m = 0; m < 128; ++m) {
                                    const int row = shift_rows + (m / 8) * p.halo_w + m % 8;
                                    if (row >= p.halo_w * p.halo_h) { printf("%s: window leaves the tile\n", g.name); return 1; }
                                    for (int c = 0; c < C; ++c)
                                        halo[m] += tile[static_cast<size_t>(row) * C + c] *
                                                   w[(static_cast<size_t>(tap.brow) + n) * C + c];
                                }
                            }
                        }
                        for (int m = 0; m < 128; ++m) { ++checked; mismatches += ref[m] != halo[m]; }
                    }
                }
    printf("%-34s groups of %d taps, halo %2d x %2d, %d stages: %lld values, %lld mismatches\n", g.name, gt, p.halo_h,
           p.halo_w, p.stages, checked, mismatches);
    return mismatches != 0;
}

int main() {
    setenv("VG_HALO", "1", 1);
    const Geo cases[] = {
        {"down k4 s2 p1 64x64", false, 2, 64, 64, 4, 2, 1},
        {"down k4 s2 p1 32x48", false, 1, 32, 48, 4, 2, 1},
        {"down k3 s1 p1 32x16", false, 2, 32, 16, 3, 1, 1},
        {"down k2 s1 p0 17x9", false, 1, 17, 9, 2, 1, 0},
        {"up   k4 s2 p1 -> 64x64", true, 2, 64, 64, 4, 2, 1},
        {"up   k4 s2 p1 -> 32x48", true, 1, 32, 48, 4, 2, 1},
        {"up   k3 s1 p1 -> 32x16", true, 2, 32, 16, 3, 1, 1},
    };
    int failures = 0;
    for (const Geo& g : cases) failures += run(g);
    // geometries the plan must refuse: 8x8 grids (no 16-row tiles), narrow channel chunks
    {
        IgemmParams p;
        std::memset(&p, 0, sizeof(p));
        p.kchunk = 64; p.c_chunks = 1; p.n_tile = 64; p.num_phases = 1; p.taps_per_phase = 16;
        down_fill_taps(p, 4, 2, 1, 3);
        if (halo_plan(p, 8, 8, 4)) { printf("8x8 grid accepted\n"); ++failures; }
        p.kchunk = 32;
        if (halo_plan(p, 16, 16, 4)) { printf("64-byte rows accepted\n"); ++failures; }
    }
    printf(failures ? "FAILED\n" : "HALO PLAN OK\n");
    return failures != 0;
}
