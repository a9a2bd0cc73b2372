// Host execution of the device code (TEST INFRASTRUCTURE): the kernel bodies of sdf_b200/csrc/sdf_kernels.cuh
// driven through the same sequence of launches as sdfb200.cu's generate_impl (one round, lists sized for the
// worst case), on the SIMT stand-in of tests/hostsim/cuda_runtime.h.  Built by tests/util_hostsim.py.
//
// Evaluator: the interpreter over a ProgramImage (what the statically compiled library runs), or -- when
// HOSTSIM_SPEC names a file holding `struct SpecEval` cut out of the source sdfb200.cu's jit_source() generated
// for one tree -- the run-time specialised straight-line evaluator with its thread-run shape (kN, kShare).
#include "sdf_kernels.cuh"

#include <memory>
#include <vector>

#ifdef HOSTSIM_SPEC
#include HOSTSIM_SPEC
typedef SpecEval EV;
#define MAKE_EV(img) SpecEval()
#else
typedef InterpEval EV;
#define MAKE_EV(img) InterpEval{&(img)}
#endif

namespace {

constexpr int kTS = 32;

bool make_image(const uint32_t* code, int n_code, const double* consts, int n_consts, ProgramImage* img) {
    memset(img, 0, sizeof(*img));
    if (n_code <= 0 || n_code > SDF_MAX_CODE || n_consts < 0 || n_consts > SDF_MAX_CONSTS) return false;
    img->n_code = (uint32_t)n_code;
    memcpy(img->code, code, sizeof(uint32_t) * (size_t)n_code);
    for (int i = 0; i < n_consts; ++i) { img->cd[i] = consts[i]; img->cf[i] = (float)consts[i]; }
    return true;
}

struct Grid {
    std::vector<double> axd;
    std::vector<float> axf;
    GridDesc g;
    Grid(const double* xs, int nx, const double* ys, int ny, const double* zs, int nz, int bs) {
        axd.insert(axd.end(), xs, xs + nx); axd.insert(axd.end(), ys, ys + ny); axd.insert(axd.end(), zs, zs + nz);
        axf.resize(axd.size());
        for (size_t i = 0; i < axd.size(); ++i) axf[i] = (float)axd[i];            // axes_to_float_kernel
        g.xs = axd.data(); g.ys = g.xs + nx; g.zs = g.ys + ny;
        g.xf = axf.data(); g.yf = g.xf + nx; g.zf = g.yf + ny;
        g.nx = nx; g.ny = ny; g.nz = nz; g.bs = bs;
        g.nbx = (nx + bs - 1) / bs; g.nby = (ny + bs - 1) / bs; g.nbz = (nz + bs - 1) / bs;
        g.ub0 = 0; g.ub1 = g.nbx * g.nby * g.nbz;
        g.sub = (bs + kTS - 1) / kTS;
    }
};

std::unique_ptr<McTablesSmem> pack_tables() {                 // as sdfb200.cu get_ctx packs them
    std::unique_ptr<McTablesSmem> t(new McTablesSmem());
    memset(t.get(), 0, sizeof(McTablesSmem));
    memcpy(t->sub_base, MC_SUB_BASE, sizeof(t->sub_base));
    memcpy(t->tile_off, MC_TILE_OFF, sizeof(t->tile_off));
    memcpy(t->sub_true, MC_SUB_TRUE, sizeof(t->sub_true));
    memcpy(t->sub_false, MC_SUB_FALSE, sizeof(t->sub_false));
    memcpy(t->sub_interior, MC_SUB_INTERIOR, sizeof(t->sub_interior));
    memcpy(t->nface, MC_NFACE, sizeof(t->nface));
    memcpy(t->faces, MC_FACES, sizeof(t->faces));
    memcpy(t->tile_ntri, MC_TILE_NTRI, sizeof(t->tile_ntri));
    memcpy(t->tile_data, MC_TILE_DATA, sizeof(t->tile_data));
    return t;
}

template <int PREC, int THREADS>
void run_tile_eval(unsigned grid, const ProgramImage& img, const GridDesc& g, float tau, const unsigned int* active,
                   const TileLists& L, Counters* ctr) {
    constexpr bool F64 = PREC == SDFB200_PREC_F64;
    hostsim::launch(grid, THREADS, sizeof(TileSmem<kTS, F64>), [&] {
        tile_eval_body<kTS, PREC, THREADS>(MAKE_EV(img), g, tau, active, 0u, 0xffffffffu, L, ctr);
    });
}

}  // namespace

extern "C" {

void hs_set_fast_ulps(double ulps) { hostsim::S.fast_ulps = ulps; }
void hs_set_schedule(int mode) { hostsim::S.schedule = mode; hostsim::S.rng = 0x9e3779b97f4a7c15ull * (unsigned long long)(mode + 1); }
unsigned long hs_switches(void) { return hostsim::S.switches; }
int hs_spec_shape(int* kn, int* kshare) { *kn = EV::kN; *kshare = EV::kShare;
#ifdef HOSTSIM_SPEC
    return 1;
#else
    return 0;
#endif
}

// SDF3.__call__ on explicit points: eval_stream_f64_body / eval_stream_f32_body<FAST>
int hs_eval_f64(const uint32_t* code, int n_code, const double* consts, int n_consts, const double* xyz, long long n,
                double* out, int grid) {
    static ProgramImage img;
    if (!make_image(code, n_code, consts, n_consts, &img)) return 1;
    hostsim::launch((unsigned)grid, 256, 0, [&] { eval_stream_f64_body(MAKE_EV(img), xyz, n, out); });
    return 0;
}

int hs_eval_f32(const uint32_t* code, int n_code, const double* consts, int n_consts, const float* xyz, long long n,
                float* out, int fast, int grid) {
    static ProgramImage img;
    if (!make_image(code, n_code, consts, n_consts, &img)) return 1;
    if (fast) hostsim::launch((unsigned)grid, STREAM_THREADS, 0, [&] { eval_stream_f32_body<true>(MAKE_EV(img), xyz, n, out); });
    else hostsim::launch((unsigned)grid, STREAM_THREADS, 0, [&] { eval_stream_f32_body<false>(MAKE_EV(img), xyz, n, out); });
    return 0;
}

// core._skip for every batch of the lattice: skip_prepass_body
int hs_skip_flags(const uint32_t* code, int n_code, const double* consts, int n_consts, const double* xs, int nx,
                  const double* ys, int ny, const double* zs, int nz, int bs, unsigned char* flags) {
    static ProgramImage img;
    if (!make_image(code, n_code, consts, n_consts, &img)) return 1;
    Grid G(xs, nx, ys, ny, zs, nz, bs);
    const int nb = G.g.ub1;
    std::vector<unsigned int> active((size_t)nb), bt((size_t)nb);
    Counters ctr;
    memset(&ctr, 0, sizeof(ctr));
    const unsigned pb = (unsigned)(((long long)nb * PREPASS_LANES + 127) / 128);
    hostsim::launch(pb, 128, 0, [&] { skip_prepass_body(MAKE_EV(img), G.g, 1, flags, active.data(), bt.data(), &ctr); });
    return 0;
}

// The whole step of sdfb200_generate (one round): pre-pass, (A) evaluate + classify, (B) float64 refinement,
// (C) extraction + batch statistics.  stats = skipped, empty, nonempty, samples f32, samples f64, active cubes.
// grids[4] = CTAs of the sign / tile, classify, refine, extract launches (any value must give the same soup).
int hs_generate(const uint32_t* code, int n_code, const double* consts, int n_consts, const double* xs, int nx,
                const double* ys, int ny, const double* zs, int nz, int bs, int sparse, int precision, double tau_,
                float* tri_out, long long cap, long long* n_tri, long long* stats, const int* grids) {
    static ProgramImage img;
    if (!make_image(code, n_code, consts, n_consts, &img)) return 1;
    const float tau = (float)tau_;
    Grid G(xs, nx, ys, ny, zs, nz, bs);
    const GridDesc& g = G.g;
    const int nb = g.ub1;
    std::vector<unsigned char> flags((size_t)nb);
    std::vector<unsigned int> active((size_t)nb), bt((size_t)nb);
    Counters ctr;
    memset(&ctr, 0, sizeof(ctr));
    const unsigned pb = (unsigned)(((long long)nb * PREPASS_LANES + 127) / 128);
    hostsim::launch(pb, 128, 0, [&] { skip_prepass_body(MAKE_EV(img), g, sparse, flags.data(), active.data(), bt.data(), &ctr); });

    const size_t tiles = (size_t)ctr.n_active * g.sub * g.sub * g.sub;
    if (tiles == 0) {
        hostsim::launch((unsigned)((nb + 127) / 128), 128, 0, [&] { batch_stats_body(g, flags.data(), bt.data(), &ctr); });
    } else {
        std::vector<float> vol(tiles * VOL_STRIDE);
        std::vector<TileMeta> meta(tiles);
        std::vector<uint2> cubes(tiles * 32768), pts(tiles * 35937);
        std::vector<unsigned char> scratch(tiles * SIGN_STRIDE + 16);
        unsigned char* scr = scratch.data() + ((16 - ((uintptr_t)scratch.data() & 15u)) & 15u);
        TileLists L;
        memset(&L, 0, sizeof(L));
        L.vol = vol.data(); L.tile_meta = meta.data(); L.cubes = cubes.data(); L.pts = pts.data();
        L.slot_cap = (unsigned)tiles; L.cube_cap = (unsigned)cubes.size(); L.pt_cap = (unsigned)pts.size();
        const bool refine = precision == SDFB200_PREC_MIXED || precision == SDFB200_PREC_EXACT;
        if (precision == SDFB200_PREC_EXACT) {
            hostsim::launch((unsigned)grids[0], SIGN_THREADS, 0, [&] {
                tile_sign_body<kTS, SIGN_THREADS>(MAKE_EV(img), g, tau, active.data(), 0u, 0xffffffffu, &ctr, scr);
            });
            constexpr int SHARE = EV::kShare < 0 ? 2 : EV::kShare;
            hostsim::launch((unsigned)grids[1], CLASSIFY_THREADS, 0, [&] {
                tile_classify_body<kTS, SHARE>(g, active.data(), 0u, 0xffffffffu, L, &ctr, scr);
            });
        } else if (precision == SDFB200_PREC_F32) run_tile_eval<SDFB200_PREC_F32, 512>((unsigned)grids[0], img, g, tau, active.data(), L, &ctr);
        else if (precision == SDFB200_PREC_MIXED) run_tile_eval<SDFB200_PREC_MIXED, 512>((unsigned)grids[0], img, g, tau, active.data(), L, &ctr);
        else run_tile_eval<SDFB200_PREC_F64, 512>((unsigned)grids[0], img, g, tau, active.data(), L, &ctr);
        if (ctr.n_cubes > L.cube_cap || ctr.n_pts > L.pt_cap || ctr.n_slots > L.slot_cap) return 2;
        if (refine)
            hostsim::launch((unsigned)grids[2], 128, 0, [&] { refine_points_body<kTS>(MAKE_EV(img), g, L, &ctr); });
        const std::unique_ptr<McTablesSmem> tables = pack_tables();
        hostsim::launch((unsigned)grids[3], EXTRACT_THREADS, 0, [&] {
            extract_body<kTS>(g, L, &ctr, bt.data(), tri_out, (unsigned long long)(tri_out ? cap : 0), nullptr, tables.get(),
                              flags.data(), nullptr, nullptr);
        });
    }
    *n_tri = (long long)ctr.tri_cursor;
    stats[0] = ctr.skipped; stats[1] = ctr.empty; stats[2] = ctr.nonempty;
    stats[3] = (long long)ctr.pts_f32; stats[4] = (long long)ctr.pts_f64; stats[5] = (long long)ctr.active_cubes;
    return 0;
}

// sdfb200_mc_volume: marching cubes + world mapping of one caller-supplied batch volume (<= 33^3 samples).  The
// candidate list (sdfb200.cu volume_candidates_kernel: every cube with mixed `v > 0` corners) is built here.
int hs_mc_volume(const float* volume, int nx, int ny, int nz, const double* xs, const double* ys, const double* zs,
                 float* tri_out, long long cap, long long* n_tri, int grid) {
    *n_tri = 0;
    if (nx < 2 || ny < 2 || nz < 2) return 0;
    Grid G(xs, nx, ys, ny, zs, nz, kTS);
    std::vector<float> vol((size_t)VOL_STRIDE);
    memcpy(vol.data(), volume, sizeof(float) * (size_t)nx * ny * nz);
    TileMeta m;
    memset(&m, 0, sizeof(m));
    m.ext = (unsigned)nx | ((unsigned)ny << 8) | ((unsigned)nz << 16);
    std::vector<uint2> cubes;
    for (int cx = 0; cx < nx - 1; ++cx)
        for (int cy = 0; cy < ny - 1; ++cy)
            for (int cz = 0; cz < nz - 1; ++cz) {
                int pos = 0;
                for (int k = 0; k < 8; ++k)
                    pos += volume[((size_t)(cx + ((k >> 2) & 1)) * ny + (cy + ((k >> 1) & 1))) * nz + (cz + (k & 1))] > 0.0f;
                if (pos != 0 && pos != 8) cubes.push_back(make_uint2(0u, (unsigned)((cx << 10) | (cy << 5) | cz)));
            }
    Counters ctr;
    memset(&ctr, 0, sizeof(ctr));
    ctr.n_cubes = (unsigned)cubes.size();
    if (cubes.empty()) cubes.push_back(make_uint2(0u, 0u));
    TileLists L;
    memset(&L, 0, sizeof(L));
    L.vol = vol.data(); L.tile_meta = &m; L.cubes = cubes.data(); L.slot_cap = 1; L.cube_cap = (unsigned)cubes.size();
    std::vector<unsigned int> bt((size_t)G.g.ub1, 0u);
    const std::unique_ptr<McTablesSmem> tables = pack_tables();
    hostsim::launch((unsigned)grid, EXTRACT_THREADS, 0, [&] {
        extract_body<kTS>(G.g, L, &ctr, bt.data(), tri_out, (unsigned long long)(tri_out ? cap : 0), nullptr, tables.get(),
                          nullptr, nullptr, nullptr);
    });
    *n_tri = (long long)ctr.tri_cursor;
    return 0;
}

// sdf/stl.py records: stl_records_body
int hs_stl_records(const float* tris, long long n, unsigned char* out) {
    if (n <= 0) return 0;
    hostsim::launch((unsigned)((n + STL_THREADS - 1) / STL_THREADS), STL_THREADS, 0, [&] { stl_records_body(tris, n, out); });
    return 0;
}

}  // extern "C"
