// Host check of csrc/fft_mixed.cuh (the same __host__ __device__ code the kernels run): every 7-smooth length up to
// 12 000 against a direct O(n^2) DFT (sampled bins), float and double.   nvcc -O2 -o fft_test.bin fft_mixed_host_test.cu && ./fft_test.bin
#include <cmath>
#include <cstdio>
#include <vector>
#include "../../intrinsictimescales.jl_b200/csrc/fft_mixed.cuh"
using namespace itjl;
template <typename T> double run(int n, int nt) {
    FftPlan plan;
    if (!fft_plan_make(n, &plan)) return -1.0;
    std::vector<Cx<T>> a(n), b(n), roots(n);
    std::vector<double> xr(n), xi(n);
    unsigned st = 12345u + n;
    for (int i = 0; i < n; ++i) {
        st = st * 1664525u + 1013904223u; xr[i] = (double)(st >> 8) / 16777216.0 - 0.5;
        st = st * 1664525u + 1013904223u; xi[i] = (double)(st >> 8) / 16777216.0 - 0.5;
        a[i] = {(T)xr[i], (T)xi[i]};
        roots[i] = {(T)cos(-2.0 * M_PI * i / n), (T)sin(-2.0 * M_PI * i / n)};
    }
    // two-level roots (what the fused kernel uses) for float, the plain table for double
    const int hi = (n + (1 << kRootLoBits) - 1) >> kRootLoBits;
    std::vector<Cx<T>> two(hi + (1 << kRootLoBits));
    for (int h = 0; h < hi; ++h) two[h] = {(T)cos(-2.0 * M_PI * (h << kRootLoBits) / n), (T)sin(-2.0 * M_PI * (h << kRootLoBits) / n)};
    for (int l = 0; l < (1 << kRootLoBits); ++l) two[hi + l] = {(T)cos(-2.0 * M_PI * l / n), (T)sin(-2.0 * M_PI * l / n)};
    const RootsTable<T> rt{roots.data()};
    const RootsTwoLevel<T> r2{two.data(), two.data() + hi};
    int len = n, s = 1;
    Cx<T>* x = a.data(); Cx<T>* y = b.data();
    for (int k = 0; k < plan.n_stages; ++k) {
        for (int tid = 0; tid < nt; ++tid) {
            if (sizeof(T) == 4) fft_stage_any<T>(plan.radix[k], x, y, n, len, s, plan.magic[k], r2, tid, nt);
            else fft_stage_any<T>(plan.radix[k], x, y, n, len, s, plan.magic[k], rt, tid, nt);
        }
        len /= plan.radix[k]; s *= plan.radix[k];
        Cx<T>* t = x; x = y; y = t;
    }
    double worst = 0.0, scale = 0.0;
    for (int kk = 0; kk < 24; ++kk) {
        const int k = (int)(((long long)kk * 7919 + 3) % n);
        double re = 0.0, im = 0.0;
        for (int i = 0; i < n; ++i) {
            const double ang = -2.0 * M_PI * (double)(((long long)i * k) % n) / n;
            re += xr[i] * cos(ang) - xi[i] * sin(ang); im += xr[i] * sin(ang) + xi[i] * cos(ang);
        }
        worst = fmax(worst, hypot(x[k].x - re, x[k].y - im));
        scale = fmax(scale, hypot(re, im));
    }
    return worst / scale;
}
int main() {
    double wf = 0.0, wd = 0.0; int count = 0;
    for (int n = 1; n <= 12000; ++n) {
        if (next_fast_len(n) != n) continue;
        if (n > 600 && (n % 7 != 0) && (n % 125 != 0) && n != 5000 && n != 10000 && n != 4096 && n != 6561) continue;   // keep the run short
        const double ef = run<float>(n, 37), ed = run<double>(n, 64);
        if (ef < 0 || ed < 0) { printf("plan failed for %d\n", n); return 1; }
        wf = fmax(wf, ef); wd = fmax(wd, ed); ++count;
        if (ef > 2e-5 || ed > 1e-12) { printf("n=%d float %.3e double %.3e\n", n, ef, ed); return 1; }
    }
    printf("%d lengths ok: worst relative error float %.2e, double %.2e\n", count, wf, wd);
    return 0;
}
