// Host check of csrc/spline_table.h (compiled and run by tests/test_spline_host.py, no GPU):
// reads one layer's filter weights, builds the quintic B-spline table and reports the largest
// deviation of spline value / derivative from the directly evaluated FP64 filter.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../mlff_distiller_b200/csrc/spline_table.h"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int hdr[2];
    float rc;
    if (std::fread(hdr, sizeof(int), 2, f) != 2 || std::fread(&rc, sizeof(float), 1, f) != 1) return 2;
    const int H = hdr[0], K = hdr[1];
    std::vector<float> centers(K), gammas(K), W1((size_t)H * K), b1(H), W2((size_t)3 * H * H), b2(3 * H);
    auto rd = [&](std::vector<float>& v) { return std::fread(v.data(), sizeof(float), v.size(), f) == v.size(); };
    if (!rd(centers) || !rd(gammas) || !rd(W1) || !rd(b1) || !rd(W2) || !rd(b2)) return 2;
    std::fclose(f);
    using namespace mlffd;
    std::vector<float> tab = build_filter_spline(H, K, rc, centers.data(), gammas.data(), W1.data(), b1.data(),
                                                 W2.data(), b2.data());
    const int R = kSplineRows, C = 3 * H, n = kSplineIntervals;
    const double h = (double)rc / n;
    double max_val = 0, max_der = 0, amp = 0, damp = 0;
    std::vector<double> hidden, v0(C), vp(C), vm(C);
    const int samples = 4001;
    for (int s = 0; s < samples; ++s) {
        const double d = 0.3 + (rc - 0.3) * s / (samples - 1);
        const double x = d / h;
        const int seg = std::min((int)x, n - 1);
        double b[6], db[6];
        quintic_basis_host(x - seg, b, db);
        filter_value_host(d, H, K, rc, centers.data(), gammas.data(), W1.data(), b1.data(), W2.data(), b2.data(), hidden, v0.data());
        const double e = 1e-5;
        const double dp = std::min(d + e, (double)rc), dm = d - e;
        filter_value_host(dp, H, K, rc, centers.data(), gammas.data(), W1.data(), b1.data(), W2.data(), b2.data(), hidden, vp.data());
        filter_value_host(dm, H, K, rc, centers.data(), gammas.data(), W1.data(), b1.data(), W2.data(), b2.data(), hidden, vm.data());
        for (int c = 0; c < C; ++c) {
            const int comp = c / H, ch = c % H, sl = ch / kSliceChannels, cc = ch % kSliceChannels;
            double val = 0, der = 0;
            for (int j = 0; j < 6; ++j) {
                const double coef = tab[(((size_t)sl * R + seg + j) * 3 + comp) * kSliceChannels + cc];
                val += b[j] * coef;
                der += db[j] * coef;
            }
            der /= h;
            const double fd = (vp[c] - vm[c]) / (dp - dm);
            max_val = std::max(max_val, std::fabs(val - v0[c]));
            if (s > 0 && s < samples - 1) max_der = std::max(max_der, std::fabs(der - fd));
            amp = std::max(amp, std::fabs(v0[c]));
            damp = std::max(damp, std::fabs(fd));
        }
    }
    std::printf("{\"max_val_err\": %.3e, \"max_der_err\": %.3e, \"amp\": %.4f, \"damp\": %.4f}\n", max_val, max_der, amp, damp);
    return 0;
}
