// Small driver over the C ABI for compute-sanitizer (memcheck / racecheck on small sizes).
//   g++ -O2 -I include tools/sanitize_small.cpp -L toyni_b200 -lntt_cuda -Wl,-rpath,$PWD/toyni_b200 -o tools/sanitize_small.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "toyni_ntt_cuda.h"
static const uint32_t P = 2013265921u;
int main() {
    if (!bb_device_ok()) { printf("no sm_100 device\n"); return 1; }
    int rc = 0;
    for (uint32_t log_n = 0; log_n <= 18; log_n += (log_n < 10 ? 1 : 2)) {
        size_t n = (size_t)1 << log_n;
        std::vector<uint32_t> h(n), back(n);
        for (size_t i = 0; i < n; i++) h[i] = (uint32_t)((i * 2654435761ull + 12345) % P);
        uint32_t* d = nullptr;
        rc |= bb_dev_alloc((void**)&d, n * 4 + 16);
        rc |= bb_h2d(d, h.data(), n * 4);
        rc |= bb_ntt_device(d, log_n, 0);
        rc |= bb_ntt_device(d, log_n, 1);
        rc |= bb_d2h(back.data(), d, n * 4);
        rc |= bb_sync();
        for (size_t i = 0; i < n; i++) if (back[i] != h[i]) { printf("roundtrip mismatch log_n=%u i=%zu\n", log_n, i); return 2; }
        // coset LDE (blowup 32 where it fits) + inverse
        if (log_n >= 5 && log_n <= 16) {
            uint32_t* o = nullptr;
            rc |= bb_dev_alloc((void**)&o, n * 4);
            rc |= bb_coset_fft_device(d, n / 32, log_n, 7, 1, o);
            rc |= bb_coset_ifft_device(o, log_n, 7, 1);
            rc |= bb_d2h(back.data(), o, n * 4);
            rc |= bb_sync();
            for (size_t i = 0; i < n; i++) if (back[i] != (i < n / 32 ? h[i] : 0u)) { printf("lde mismatch log_n=%u i=%zu\n", log_n, i); return 3; }
            bb_dev_free(o);
        }
        bb_dev_free(d);
    }
    // fold + commit loop on 2^10 Ext values, betas up front, with hashing
    {
        size_t n = 1 << 10;
        std::vector<uint32_t> h(4 * n);
        for (size_t i = 0; i < 4 * n; i++) h[i] = (uint32_t)((i * 40503ull + 7) % P);
        size_t nsalt = 0, nnodes = 0, tot = 0;
        for (size_t m = n; m >= 16; m /= 2) { nnodes += bb_merkle_node_count(m); if (m > 16) nsalt += m; if (m < n) tot += m; }
        uint32_t *l0, *layers; uint8_t *salts, *nodes;
        rc |= bb_dev_alloc((void**)&l0, 16 * n); rc |= bb_dev_alloc((void**)&layers, 16 * tot);
        rc |= bb_dev_alloc((void**)&salts, 16 * nsalt); rc |= bb_dev_alloc((void**)&nodes, 32 * nnodes);
        std::vector<uint8_t> hs(16 * nsalt, 0x5a);
        rc |= bb_h2d(l0, h.data(), 16 * n); rc |= bb_h2d(salts, hs.data(), hs.size());
        std::vector<uint32_t> betas(4 * 6, 12345);
        std::vector<uint8_t> roots(32 * 7);
        size_t folds = 0;
        rc |= bb_fri_commit_device(l0, n, 7, 16, 4, salts, nullptr, nullptr, betas.data(), layers, nodes, roots.data(), &folds);
        rc |= bb_sync();
        printf("commit loop folds=%zu root[0]=%02x%02x\n", folds, roots[0], roots[1]);
    }
    printf("rc=%d last_error=%d\n", rc, bb_last_error());
    return rc ? 4 : 0;
}
