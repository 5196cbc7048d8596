// Sanitizer harness for the host-side BVH passes of csrc/lbvh.cu (SAH build / rebuild, re-insertion, BVH4 collapse): random reference
// sets (scattered, coplanar, identical, overlapping), then bit-flipped node arrays that must be rejected or survived without a crash.
// Not part of the pytest run (it needs ~30 s of nvcc); build and run by hand:
//   S=vulkan-path-tracer_b200/csrc; F="-Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined"
//   nvcc -std=c++17 -O1 -g -gencode arch=compute_100a,code=sm_100a $F -I$S -Iinclude -c $S/lbvh.cu -o /tmp/lbvh_san.o
//   nvcc -std=c++17 -O1 -g -gencode arch=compute_100a,code=sm_100a $F -I$S -Iinclude tests/native/fuzz_bvh_passes.cpp /tmp/lbvh_san.o -o /tmp/fuzz_bvh -lcudart
//   ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0 /tmp/fuzz_bvh        (no GPU needed: only host functions are called)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include <cuda_runtime.h>
#include "kernels.h"
using namespace b200pt;
int main() {
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    for (int iter = 0; iter < 300; iter++) {
        uint32_t n = 5 + rng() % (iter < 250 ? 400 : (iter < 290 ? 20000 : 150000));   // the last ten are large enough for the thread pool of the rebuild
        std::vector<float> rb((size_t)n * 6);
        const int mode = iter % 4;
        for (uint32_t i = 0; i < n; i++) {
            float c[3] = { U(rng), U(rng), U(rng) };
            if (mode == 1) { c[0] = 0.5f; }                   // coplanar centroids
            if (mode == 2) { c[0] = c[1] = c[2] = 0.25f; }    // all identical
            float e = mode == 3 ? U(rng) * 0.5f : 0.01f * U(rng);
            for (int a = 0; a < 3; a++) { rb[i * 6 + a] = c[a] - e; rb[i * 6 + 3 + a] = c[a] + e; }
        }
        std::vector<BvhNode> nodes(n), out(n), out2(n), out3(n); std::vector<uint32_t> perm(n); int d = 0; double c = 0, s[2];
        uint32_t m = bvh2_sah_build_host(rb.data(), n, 1.0f + (iter % 3), nodes.data(), perm.data(), &d, &c);
        if (m == 0 || m > n - 1) { printf("build failed n=%u\n", n); return 1; }
        uint32_t m2 = bvh2_sah_rebuild_host(nodes.data(), m, 0, out.data(), &d, s);
        if (m2 != m) { printf("rebuild count %u != %u\n", m2, m); return 1; }
        uint32_t m3 = bvh2_reinsert_host(out.data(), m2, 0, out2.data(), 1 + iter % 3, 0.1f + 0.3f * (iter % 3), &d, s);
        if (m3 != m && m >= 2) { printf("reinsert count %u != %u\n", m3, m); return 1; } if (m3 == 0) continue;
        if (s[1] > s[0] * 1.00001) { printf("cost went up %g -> %g\n", s[0], s[1]); return 1; }
        std::vector<Bvh4Node> n4(m); int d4 = 0;
        uint32_t m4 = bvh4_collapse_host(out2.data(), m3, 0, n4.data(), &d4);
        if (m4 == 0) { printf("collapse failed\n"); return 1; }
        // corrupt inputs must not crash
        std::vector<BvhNode> bad = out2;
        for (int k = 0; k < 8; k++) { uint32_t w = rng() % (m3 * 16); reinterpret_cast<uint32_t *>(bad.data())[w] ^= (1u << (rng() % 32)); }
        bvh2_sah_rebuild_host(bad.data(), m3, 0, out3.data(), &d, s);
        bvh2_reinsert_host(bad.data(), m3, 0, out3.data(), 2, 0.3f, &d, s);
        bvh4_collapse_host(bad.data(), m3, 0, n4.data(), &d4);
    }
    puts("ok");
    return 0;
}
