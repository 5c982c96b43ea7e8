// tools/host_prepare_bench.cpp - times (and hashes) every phase of the host-side scene preparation on the n x n terrain:
// what crt_build_bvh does for the caller and what crt_create derives (csrc/host_scene.cpp).  No GPU, no CUDA:
//   bash tools/host_prepare_bench.sh 707          # 999 700 triangles; minima over REPS=6 runs
// The hashes make two builds of host_scene.cpp comparable byte for byte (git stash / another checkout via CSRC=<dir>).
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

#include "host_scene.h"
using namespace crt;

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
    const uint8_t* b = (const uint8_t*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
template <class T>
static std::vector<T> read_all(const std::string& f) {
    FILE* fp = fopen(f.c_str(), "rb");
    if (!fp) { perror(f.c_str()); exit(1); }
    fseek(fp, 0, SEEK_END);
    const size_t n = (size_t)ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<T> v(n / sizeof(T));
    if (fread(v.data(), 1, n, fp) != n) exit(1);
    fclose(fp);
    return v;
}
static std::map<std::string, double> best;
static std::vector<std::string> order;
static void phase(const char* name, double seconds, bool print, const std::string& what) {
    if (!best.count(name)) { best[name] = seconds; order.push_back(name); } else best[name] = std::min(best[name], seconds);
    if (print) printf("%-22s %.3f s  %s\n", name, seconds, what.c_str());
}
static std::string hex(uint64_t v) { char b[32]; snprintf(b, sizeof b, "%016lx", (unsigned long)v); return b; }

// the record vectors' type differs between checkouts (std::vector / RawVec): deduced from make_tri_records itself
template <class A, class B>
static void time_records(void (*fn)(const RefTriangle*, uint32_t, const RefNode*, uint32_t, A&, B&), const std::vector<RefTriangle>& tris,
                         const std::vector<RefNode>& nodes, bool print) {
    A isect;
    B shade;
    const double t0 = now();
    fn(tris.data(), (uint32_t)tris.size(), nodes.data(), (uint32_t)nodes.size(), isect, shade);
    const double dt = now() - t0;
    phase("make_tri_records", dt, print, hex(fnv(isect.data(), isect.size() * 48)) + " " + hex(fnv(shade.data(), shade.size() * 80)));
}

static void run(const std::string& dir, bool print) {
    auto tris = read_all<RefTriangle>(dir + "/pre_tris.bin");            // OBJ order
    const auto pos = read_all<float>(dir + "/positions.bin");
    const uint32_t nt = (uint32_t)tris.size();
    std::vector<RefNode> nodes;
    double dt, t0 = now();
    build_reference_bvh(tris.data(), pos.data(), nt, nodes);              // reorders tris into BVH order
    dt = now() - t0;
    phase("build_reference_bvh", dt, print, "nodes " + std::to_string(nodes.size()) + " " + hex(fnv(nodes.data(), nodes.size() * 32)) + " " + hex(fnv(tris.data(), tris.size() * 128)));
    // the region crt_create would use: every vertex and the terrain camera's eye, grown by 5 % of the diagonal
    F3 lo{FLT_MAX, FLT_MAX, FLT_MAX}, hi{-FLT_MAX, -FLT_MAX, -FLT_MAX};
    auto grow = [&](double x, double y, double z) {
        lo = {std::min(lo.x, (float)x), std::min(lo.y, (float)y), std::min(lo.z, (float)z)};
        hi = {std::max(hi.x, (float)x), std::max(hi.y, (float)y), std::max(hi.z, (float)z)};
    };
    for (const RefTriangle& t : tris) {
        grow(t.v1.x, t.v1.y, t.v1.z);
        grow((double)t.v1.x + t.e1.x, (double)t.v1.y + t.e1.y, (double)t.v1.z + t.e1.z);
        grow((double)t.v1.x + t.e2.x, (double)t.v1.y + t.e2.y, (double)t.v1.z + t.e2.z);
    }
    grow(0, 12, 40);
    const double dx = (double)hi.x - lo.x, dy = (double)hi.y - lo.y, dz = (double)hi.z - lo.z, m = 0.05 * std::sqrt(dx * dx + dy * dy + dz * dz) + 1e-3;
    lo = {(float)(lo.x - m), (float)(lo.y - m), (float)(lo.z - m)};
    hi = {(float)(hi.x + m), (float)(hi.y + m), (float)(hi.z + m)};

    t0 = now();
    std::vector<RefNode> clustered;
    recluster_leaves(nodes.data(), (uint32_t)nodes.size(), clustered);
    dt = now() - t0;
    phase("recluster_leaves", dt, print, hex(fnv(clustered.data(), clustered.size() * 32)));
    std::vector<WideNode> wide;
    t0 = now();
    const WideStats ws = flatten_wide(clustered.data(), (uint32_t)clustered.size(), tris.data(), nt, wide);
    dt = now() - t0;
    phase("flatten_wide (ref)", dt, print, "wide " + std::to_string(ws.n_wide) + " depth " + std::to_string(ws.depth) + " " + hex(fnv(wide.data(), wide.size() * 128)));
    time_records(&make_tri_records, tris, nodes, print);
    TightBvh tb;
    t0 = now();
    build_tight_bvh(tris.data(), nt, nodes.data(), (uint32_t)nodes.size(), lo, hi, kPadAuto, tb);
    dt = now() - t0;
    phase("build_tight_bvh", dt, print, "nodes " + std::to_string(tb.nodes.size()) + " split " + std::to_string(tb.n_split) + (tb.dynamic_pad ? " dynamic " : " static ") +
          hex(fnv(tb.nodes.data(), tb.nodes.size() * 32)) + " " + hex(fnv(tb.records.data(), tb.records.size() * 80)) + " " + hex(fnv(tb.boxes.data(), tb.boxes.size() * 4)));
    t0 = now();
    const SahCost a = sah_cost(tb.nodes.data(), (uint32_t)tb.nodes.size(), lo, hi), b = sah_cost(clustered.data(), (uint32_t)clustered.size(), lo, hi);
    dt = now() - t0;
    phase("sah_cost x 2", dt, print, std::to_string(a.nodes) + " " + std::to_string(a.tris) + " / " + std::to_string(b.nodes) + " " + std::to_string(b.tris));
    std::vector<WideNode> tw;
    t0 = now();
    const WideStats ts = flatten_wide(tb.nodes.data(), (uint32_t)tb.nodes.size(), tb.tris.data(), nt, tw);
    dt = now() - t0;
    phase("flatten_wide (tight)", dt, print, "wide " + std::to_string(ts.n_wide) + " depth " + std::to_string(ts.depth) + " " + hex(fnv(tw.data(), tw.size() * 128)));
}

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : ".";
    const int reps = getenv("REPS") ? atoi(getenv("REPS")) : 6;
    for (int r = 0; r < reps; r++) run(dir, r == 0);
    printf("minimum of %d runs, %u hardware threads:\n", reps, std::thread::hardware_concurrency());
    for (const std::string& k : order) printf("  %-22s %.3f s\n", k.c_str(), best[k]);
    return 0;
}
