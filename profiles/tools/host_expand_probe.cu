// Probe: can host threads expand a zero-chunk-compacted observation stream faster than PCIe delivers the dense one?
#include <cuda_runtime.h>
#include <immintrin.h>
#include <pthread.h>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

constexpr int kChunksPerEnv = 388;           // 6208 B / 16
constexpr int kWords = 13;                   // bitmap words per env
struct Job { const uint32_t* bitmap; const __m128i* src; __m128i* dst; int envs; bool nt; };

static void expand(const Job& j) {
    const __m128i zero = _mm_setzero_si128();
    const __m128i* s = j.src;
    for (int e = 0; e < j.envs; ++e) {
        const uint32_t* bm = j.bitmap + (size_t)e * kWords;
        __m128i* d = j.dst + (size_t)e * kChunksPerEnv;
        for (int w = 0; w < kWords; ++w) {
            const uint32_t m = bm[w];
            const int n = (w == kWords - 1) ? kChunksPerEnv - 32 * w : 32;
            __m128i* dw = d + 32 * w;
            if (m == 0u) {
                for (int b = 0; b < n; ++b) _mm_stream_si128(dw + b, zero);
            } else {
                // branchless: always load the next compact chunk, keep it if the bit is set
                for (int b = 0; b < n; ++b) {
                    const uint32_t bit = (m >> b) & 1u;
                    const __m128i keep = _mm_set1_epi32(-(int)bit);
                    const __m128i v = _mm_and_si128(_mm_loadu_si128(s), keep);
                    s += bit;
                    if (j.nt) _mm_stream_si128(dw + b, v); else _mm_store_si128(dw + b, v);
                }
            }
        }
    }
    _mm_sfence();
}

int main(int argc, char** argv) {
    const int B = 65536;
    const size_t dense_bytes = (size_t)B * kChunksPerEnv * 16;
    printf("host threads available: %u\n", std::thread::hardware_concurrency());
    uint8_t *dense, *compact; uint32_t* bitmap; uint8_t* dev;
    CK(cudaMallocHost(&dense, dense_bytes));
    CK(cudaMallocHost(&compact, dense_bytes));
    CK(cudaMallocHost(&bitmap, (size_t)B * kWords * 4));
    CK(cudaMalloc(&dev, dense_bytes));
    CK(cudaMemset(dev, 1, dense_bytes));
    memset(dense, 0, dense_bytes); memset(compact, 1, dense_bytes);
    // 45 % of the chunks non-zero, pseudo-random
    std::vector<size_t> env_off(B + 1, 0);
    uint64_t s = 88172645463325252ull; size_t nz = 0; int run = 0; bool on = false;
    for (int e = 0; e < B; ++e) {
        env_off[e] = nz;
        for (int w = 0; w < kWords; ++w) {
            uint32_t m = 0;
            const int n = (w == kWords - 1) ? kChunksPerEnv - 32 * w : 32;
            for (int b = 0; b < n; ++b) { if (run == 0) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; on = (s % 100) < 45; run = 1 + (int)((s >> 20) % 5); } --run; if (on) { m |= 1u << b; ++nz; } }
            bitmap[(size_t)e * kWords + w] = m;
        }
    }
    env_off[B] = nz;
    const size_t compact_bytes = nz * 16;
    printf("dense %.1f MB, compact %.1f MB\n", dense_bytes / 1e6, compact_bytes / 1e6);
    cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    auto d2h = [&](size_t bytes, uint8_t* dst) {
        double best = 1e9;
        for (int r = 0; r < 5; ++r) { double t = now(); CK(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); best = std::min(best, now() - t); }
        return best;
    };
    double t_dense = d2h(dense_bytes, dense), t_comp = d2h(compact_bytes, compact);
    printf("D2H dense %.3f ms (%.1f GB/s), compact %.3f ms (%.1f GB/s)\n", t_dense * 1e3, dense_bytes / t_dense / 1e9, t_comp * 1e3, compact_bytes / t_comp / 1e9);
    const int pieces = 128; const int envs_per_piece = B / pieces;
    for (int nt = 1; nt < 2; ++nt)
        for (int T : {1, 8, 12, 14, 15, 16}) {
            if (T > (int)std::thread::hardware_concurrency()) continue;
            for (int with_dma = 0; with_dma < 2; ++with_dma) {
                double best = 1e9, dma_rate = 0;
                for (int r = 0; r < 4; ++r) {
                    std::atomic<int> next{0}; std::atomic<bool> stop{false};
                    std::atomic<long long> dma_bytes{0};
                    std::thread dma;
                    if (with_dma) dma = std::thread([&] {
                        while (!stop.load()) { CK(cudaMemcpyAsync(compact + compact_bytes, dev, 64 << 20, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); dma_bytes += 64 << 20; }
                    });
                    if (with_dma) std::this_thread::sleep_for(std::chrono::milliseconds(2));
                    const long long dma0 = dma_bytes.load();
                    double t = now();
                    std::vector<std::thread> th;
                    for (int i = 0; i < T; ++i) th.emplace_back([&] {
                        for (;;) { int p = next.fetch_add(1); if (p >= pieces) break;
                            Job j{bitmap + (size_t)p * envs_per_piece * kWords, (const __m128i*)compact + env_off[(size_t)p * envs_per_piece],
                                  (__m128i*)dense + (size_t)p * envs_per_piece * kChunksPerEnv, envs_per_piece, nt != 0};
                            expand(j); }
                    });
                    for (auto& x : th) x.join();
                    double dt = now() - t;
                    const long long dma1 = dma_bytes.load();
                    stop = true; if (with_dma) dma.join();
                    if (dt < best) { best = dt; dma_rate = (dma1 - dma0) / dt / 1e9; }
                }
                printf("expand %s stores, %2d threads%s: %.3f ms (dense %.1f GB/s)%s", nt ? "NT" : "regular", T, with_dma ? " + concurrent D2H" : "", best * 1e3, dense_bytes / best / 1e9, with_dma ? "" : "\n");
                if (with_dma) printf(", DMA meanwhile ~%.1f GB/s\n", dma_rate);
            }
        }
    return 0;
}
