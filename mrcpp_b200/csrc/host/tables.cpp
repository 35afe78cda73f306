// Numerical tables: two-scale filters, cross-correlation coefficients, Gauss-Legendre quadrature,
// interpolating scaling functions, Hilbert-curve state tables.
#include "mrx_host.hpp"

#include <dlfcn.h>

#include <cstring>
#include <map>
#include <mutex>

namespace mrx {

void abort_msg(const char *file, int line, const std::string &msg) {
    // reference error convention: print + abort (src/utils/Printer.h:165-169)
    std::fprintf(stderr, "mrx abort %s:%d: %s\n", file, line, msg.c_str());
    std::abort();
}

namespace {
std::string g_table_path;
std::mutex g_mutex;

struct RawTables {
    // kind -> order -> data
    std::map<int, std::vector<double>> t[9];
    bool loaded = false;
};
RawTables g_raw;

void load_raw() {
    if (g_raw.loaded) return;
    if (g_table_path.empty()) {
        const char *env = std::getenv("MRX_TABLES");
        if (env) g_table_path = env;
    }
    if (g_table_path.empty()) {
        // next to the library: <dir of libmrcpp_b200.so>/../data/mwtables.bin (the in-tree layout), like the reference falls
        // back to its install directory (details::find_filters, src/utils/details.cpp:53-69)
        Dl_info info;
        if (dladdr(reinterpret_cast<const void *>(&set_table_path), &info) && info.dli_fname) {
            std::string lib(info.dli_fname);
            const size_t slash = lib.rfind('/');
            const std::string dir = slash == std::string::npos ? std::string(".") : lib.substr(0, slash);
            const std::string cand = dir + "/../data/mwtables.bin";
            if (FILE *probe = std::fopen(cand.c_str(), "rb")) {
                std::fclose(probe);
                g_table_path = cand;
            }
        }
    }
    if (g_table_path.empty()) MRX_ABORT("table path not set (mrx_init / MRX_TABLES) and no data/mwtables.bin next to the library");
    FILE *f = std::fopen(g_table_path.c_str(), "rb");
    if (!f) MRX_ABORT("cannot open table file " + g_table_path);
    char magic[4];
    int32_t n = 0;
    if (std::fread(magic, 1, 4, f) != 4 || std::memcmp(magic, "MRXT", 4) != 0) MRX_ABORT("bad table magic");
    if (std::fread(&n, 4, 1, f) != 1) MRX_ABORT("bad table header");
    for (int e = 0; e < n; e++) {
        int32_t hdr[3];
        if (std::fread(hdr, 4, 3, f) != 3) MRX_ABORT("truncated table file");
        std::vector<double> d(hdr[2]);
        if (std::fread(d.data(), 8, d.size(), f) != d.size()) MRX_ABORT("truncated table file");
        if (hdr[0] < 0 || hdr[0] > 8) MRX_ABORT("bad table kind");
        g_raw.t[hdr[0]][hdr[1]] = std::move(d);
    }
    std::fclose(f);
    g_raw.loaded = true;
}

std::map<int, std::unique_ptr<FilterSet>> g_filters;
std::map<int, std::unique_ptr<CrossCorr>> g_cc;
std::map<int, std::unique_ptr<Quadrature>> g_quad;
} // namespace

void set_table_path(const std::string &path) {
    std::lock_guard<std::mutex> lock(g_mutex);
    g_table_path = path;
}
const std::string &table_path() { return g_table_path; }

const FilterSet &filter_set(int k) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_filters.find(k);
    if (it != g_filters.end()) return *it->second;
    load_raw();
    if (!g_raw.t[0].count(k) || !g_raw.t[1].count(k)) MRX_ABORT("no filter table for order " + std::to_string(k));
    auto fs = std::make_unique<FilterSet>();
    int K = k + 1;
    fs->k = k;
    fs->K = K;
    fs->H0 = g_raw.t[0][k];
    fs->G0 = g_raw.t[1][k];
    fs->H1.assign(K * K, 0.0);
    fs->G1.assign(K * K, 0.0);
    // interpolating symmetry, MWFilter.cpp:229-235
    for (int i = 0; i < K; i++)
        for (int j = 0; j < K; j++) {
            fs->G1[i * K + j] = std::pow(-1.0, i + K) * fs->G0[i * K + (K - j - 1)];
            fs->H1[i * K + j] = fs->H0[(K - i - 1) * K + (K - j - 1)];
        }
    auto transpose = [K](const std::vector<double> &a) {
        std::vector<double> t(K * K);
        for (int i = 0; i < K; i++)
            for (int j = 0; j < K; j++) t[j * K + i] = a[i * K + j];
        return t;
    };
    // MWFilter::getSubFilter, MWFilter.cpp:100-133
    fs->sub[Compression][0] = transpose(fs->H0);
    fs->sub[Compression][1] = transpose(fs->H1);
    fs->sub[Compression][2] = transpose(fs->G0);
    fs->sub[Compression][3] = transpose(fs->G1);
    fs->sub[Reconstruction][0] = fs->H0;
    fs->sub[Reconstruction][1] = fs->G0;
    fs->sub[Reconstruction][2] = fs->H1;
    fs->sub[Reconstruction][3] = fs->G1;
    auto *p = fs.get();
    g_filters[k] = std::move(fs);
    return *p;
}

const std::vector<double> &derivative_table(int kind, int k) {
    std::lock_guard<std::mutex> lock(g_mutex);
    load_raw();
    if (kind < 4 || kind > 8) MRX_ABORT("derivative_table: bad kind");
    auto it = g_raw.t[kind].find(k);
    if (it == g_raw.t[kind].end()) MRX_ABORT("Scaling order not supported");
    if ((int)it->second.size() != 3 * (k + 1) * (k + 1)) MRX_ABORT("derivative_table: bad size");
    return it->second;
}

const CrossCorr &cross_corr(int k) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_cc.find(k);
    if (it != g_cc.end()) return *it->second;
    load_raw();
    if (!g_raw.t[2].count(k) || !g_raw.t[3].count(k)) MRX_ABORT("no cross-correlation table for order " + std::to_string(k));
    auto cc = std::make_unique<CrossCorr>();
    cc->k = k;
    cc->K = k + 1;
    cc->L = g_raw.t[2][k];
    cc->R = g_raw.t[3][k];
    for (auto &v : cc->L)
        if (std::abs(v) < MachinePrec) v = 0.0;
    for (auto &v : cc->R)
        if (std::abs(v) < MachinePrec) v = 0.0;
    auto *p = cc.get();
    g_cc[k] = std::move(cc);
    return *p;
}

namespace {
// Legendre P_n(z) and derivative on [-1,1] by the three-term recurrence.
void legendre(int n, double z, double &p, double &dp) {
    double p0 = 1.0, p1 = z;
    if (n == 0) {
        p = 1.0;
        dp = 0.0;
        return;
    }
    for (int j = 2; j <= n; j++) {
        double pj = ((2.0 * j - 1.0) * z * p1 - (j - 1.0) * p0) / j;
        p0 = p1;
        p1 = pj;
    }
    p = p1;
    dp = n * (z * p1 - p0) / (z * z - 1.0);
}
} // namespace

const Quadrature &quadrature(int n) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_quad.find(n);
    if (it != g_quad.end()) return *it->second;
    auto q = std::make_unique<Quadrature>();
    q->n = n;
    q->roots.assign(n, 0.0);
    q->weights.assign(n, 0.0);
    // GaussQuadrature::calcGaussPtsWgts (GaussQuadrature.cpp:157-193): Newton on P_n, EPS=3e-12, <=10 iterations
    int Kh = (n % 2 == 0) ? n / 2 : (n + 1) / 2;
    std::vector<double> ur(n), uw(n);
    for (int i = 0; i < Kh; i++) {
        double z = std::cos(pi * (i + 0.75) / (n + 0.5));
        double p = 0, dp = 0;
        int iter;
        for (iter = 0; iter < 10; iter++) {
            legendre(n, z, p, dp);
            double z1 = z;
            z = z1 - p / dp;
            if (std::abs(z - z1) <= 3.0e-12) break;
        }
        if (iter == 10) MRX_ABORT("Gauss-Legendre Newton iteration failed");
        ur[i] = -z;
        ur[n - 1 - i] = z;
        uw[i] = 2.0 / ((1.0 - z * z) * dp * dp);
        uw[n - 1 - i] = uw[i];
    }
    // scaled to [0,1] (calcScaledPtsWgts, GaussQuadrature.cpp:134-150)
    for (int j = 0; j < n; j++) {
        q->roots[j] = ur[j] * 0.5 + 0.0 + 0.5;
        q->weights[j] = uw[j] * 0.5;
    }
    auto *p = q.get();
    g_quad[n] = std::move(q);
    return *p;
}

// phi_j(x) = sqrt(w_j) sum_i (2i+1) P_i(2x_j-1) P_i(2x-1)   (InterpolatingBasis.cpp:62-84)
double interp_scaling_eval(int k, int j, double x) {
    const Quadrature &q = quadrature(k + 1);
    double zj = 2.0 * q.roots[j] - 1.0, z = 2.0 * x - 1.0;
    double sum = 0.0;
    for (int i = 0; i <= k; i++) {
        double a, da, b, db;
        legendre(i, zj, a, da);
        if (std::abs(std::abs(z) - 1.0) < 1e-300) {
            b = (z > 0 || i % 2 == 0) ? 1.0 : -1.0;
        } else {
            legendre(i, z, b, db);
        }
        sum += (2.0 * i + 1.0) * a * b;
    }
    return std::sqrt(q.weights[j]) * sum;
}

double interp_scaling_deriv(int k, int j, double x) {
    const Quadrature &q = quadrature(k + 1);
    double zj = 2.0 * q.roots[j] - 1.0, z = 2.0 * x - 1.0;
    double sum = 0.0;
    for (int i = 0; i <= k; i++) {
        double a, da, b, db;
        legendre(i, zj, a, da);
        legendre(i, z, b, db);
        sum += (2.0 * i + 1.0) * a * db * 2.0; // d/dx P_i(2x-1) = 2 P_i'(z)
    }
    return std::sqrt(q.weights[j]) * sum;
}

// ------------------------------------------------------------------ Hilbert tables
// State machine of the 3-D / 2-D Hilbert curve used for child traversal order
// (values as in src/trees/HilbertPath.cpp:30-101; hTable is the inverse permutation of zTable).
namespace {
const int8_t pTable2[4][4] = {{1, 0, 0, 3}, {0, 1, 1, 2}, {3, 2, 2, 1}, {2, 3, 3, 0}};
const int8_t zTable2[4][4] = {{0, 2, 3, 1}, {0, 1, 3, 2}, {3, 1, 0, 2}, {3, 2, 0, 1}};
const int8_t pTable3[12][8] = {{1, 2, 2, 9, 9, 8, 8, 4},   {2, 0, 0, 7, 7, 3, 3, 11},  {0, 1, 1, 5, 5, 10, 10, 6},
                               {4, 5, 5, 6, 6, 11, 11, 1}, {5, 3, 3, 10, 10, 0, 0, 8}, {3, 4, 4, 2, 2, 7, 7, 9},
                               {7, 8, 8, 3, 3, 2, 2, 10},  {8, 6, 6, 1, 1, 9, 9, 5},   {6, 7, 7, 11, 11, 4, 4, 0},
                               {10, 11, 11, 0, 0, 5, 5, 7}, {11, 9, 9, 4, 4, 6, 6, 2}, {9, 10, 10, 8, 8, 1, 1, 3}};
const int8_t zTable3[12][8] = {{0, 2, 6, 4, 5, 7, 3, 1}, {0, 4, 5, 1, 3, 7, 6, 2}, {0, 1, 3, 2, 6, 7, 5, 4},
                               {3, 1, 5, 7, 6, 4, 0, 2}, {3, 7, 6, 2, 0, 4, 5, 1}, {3, 1, 0, 2, 6, 4, 5, 7},
                               {5, 7, 3, 1, 0, 2, 6, 4}, {5, 1, 0, 4, 6, 2, 3, 7}, {5, 4, 6, 7, 3, 2, 0, 1},
                               {6, 4, 0, 2, 3, 1, 5, 7}, {6, 2, 3, 7, 5, 1, 0, 4}, {6, 7, 5, 4, 0, 1, 3, 2}};
} // namespace

int hilbert_child_path(int D, int path, int hIdx) {
    if (D == 1) return 0;
    if (D == 2) return pTable2[path][hIdx];
    return pTable3[path][hIdx];
}
int hilbert_z_index(int D, int path, int hIdx) {
    if (D == 1) return hIdx;
    if (D == 2) return zTable2[path][hIdx];
    return zTable3[path][hIdx];
}
int hilbert_h_index(int D, int path, int zIdx) {
    int n = 1 << D;
    for (int h = 0; h < n; h++)
        if (hilbert_z_index(D, path, h) == zIdx) return h;
    return -1;
}

} // namespace mrx
