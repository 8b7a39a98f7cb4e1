// CPU-only check of the host side of mate_b200_step_host's compacted legs (mate_b200/csrc/mate_hostpath.cuh): the tables and
// streams the two device kernels would produce are built here on the host (same layout, blocks placed in the stream in a
// shuffled order like the kernels' atomicAdd does), expand_blocks / ExpandPool rebuild or patch the rows, and the result is
// compared byte by byte.  No CUDA call is made: runs without a GPU (tests/test_hostpath_cpu.py builds it with nvcc).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../mate_b200/csrc/mate_hostpath.cuh"

using namespace mate;

struct Chunk { uint32_t w[4]; };
static bool is_zero(const Chunk& c) { return (c.w[0] | c.w[1] | c.w[2] | c.w[3]) == 0u; }
static bool same(const Chunk& a, const Chunk& b) { return memcmp(&a, &b, 16) == 0; }

// what compact_chunks_kernel (changes == nullptr) / compact_changes_kernel leave behind
static void compact(const std::vector<Chunk>& rows, const std::vector<Chunk>* before, std::vector<CompactEntry>& table,
                    std::vector<Chunk>& stream, std::mt19937& rng) {
    const long long nchunks = (long long)rows.size(), nblocks = (nchunks + kCompactBlock - 1) / kCompactBlock;
    table.assign(nblocks, CompactEntry{});
    std::vector<long long> order(nblocks);
    for (long long b = 0; b < nblocks; ++b) order[b] = b;
    std::shuffle(order.begin(), order.end(), rng);
    stream.clear();
    for (long long b : order) {
        CompactEntry e{};
        e.offset = (uint32_t)stream.size();
        for (int w = 0; w < kCompactWords; ++w) {
            uint32_t bits = 0;
            for (int k = 0; k < 32; ++k) {
                const long long i = b * kCompactBlock + 32 * w + k;
                if (i >= nchunks) break;
                const bool keep = before ? !same(rows[i], (*before)[i]) : !is_zero(rows[i]);
                bits |= (uint32_t)keep << k;
            }
            if (before) { bits |= bits >> 1; bits |= bits >> 2; bits = (bits & 0x11111111u) * 0xFu; }
            e.words[w] = bits;
            for (int k = 0; k < 32; ++k)
                if ((bits >> k) & 1u) stream.push_back(rows[b * kCompactBlock + 32 * w + k]);
        }
        table[b] = e;
    }
    for (int pad = 0; pad < 8; ++pad) stream.push_back(Chunk{{0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu}});
}

static std::vector<Chunk> random_rows(long long n, double p_zero, std::mt19937& rng) {
    std::vector<Chunk> rows(n);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    bool zero_run = false;
    for (auto& c : rows) {
        if (u(rng) < 0.3) zero_run = u(rng) < p_zero;     // runs of zeros and of data, like entity slots
        for (auto& w : c.w) w = zero_run ? 0u : (uint32_t)rng() | 1u;
    }
    return rows;
}

int main() {
    std::mt19937 rng(12345);
    int failures = 0;
    ExpandPool pool(3, 0);
    for (const long long nchunks : {4LL, 252LL, 256LL, 260LL, 1000LL, 4096LL, 100004LL}) {
        for (const int misalign : {0, 1}) {                        // destination 16-byte aligned or only 4-byte aligned
            for (const double p_zero : {0.0, 0.55, 1.0}) {
                const std::vector<Chunk> before = random_rows(nchunks, p_zero, rng);
                std::vector<Chunk> now = before;
                std::uniform_real_distribution<double> u(0.0, 1.0);
                for (auto& c : now) if (u(rng) < 0.2) { for (auto& w : c.w) w = u(rng) < 0.3 ? 0u : (uint32_t)rng(); }
                std::vector<CompactEntry> table;
                std::vector<Chunk> stream;
                std::vector<unsigned char> raw((size_t)nchunks * 16 + 64);
                unsigned char* base = raw.data() + ((64 - ((uintptr_t)raw.data() & 63)) & 63) + 4 * misalign;
                const long long nblocks = (nchunks + kCompactBlock - 1) / kCompactBlock;
                auto run = [&](bool only_marked) {
                    const long long per_piece = std::max(1LL, (nblocks + 4) / 5);
                    for (long long b0 = 0; b0 < nblocks; b0 += per_piece)
                        pool.submit(ExpandPool::Work{table.data(), only_marked, reinterpret_cast<const __m128i*>(stream.data()),
                                                     reinterpret_cast<__m128i*>(base), nchunks, b0, std::min(nblocks, b0 + per_piece), misalign == 0, nullptr});
                    pool.wait();
                };
                // leg 1: non-zero chunks kept, every byte rebuilt (the buffer holds garbage before)
                compact(now, nullptr, table, stream, rng);
                memset(base, 0xA5, (size_t)nchunks * 16);
                run(false);
                if (memcmp(base, now.data(), (size_t)nchunks * 16) != 0) { ++failures; printf("rebuild mismatch: n=%lld misalign=%d p=%.2f\n", nchunks, misalign, p_zero); }
                // leg 2: the buffer holds the previous rows, the changed 64-byte groups are patched in
                compact(now, &before, table, stream, rng);
                memcpy(base, before.data(), (size_t)nchunks * 16);
                run(true);
                if (memcmp(base, now.data(), (size_t)nchunks * 16) != 0) { ++failures; printf("patch mismatch: n=%lld misalign=%d p=%.2f\n", nchunks, misalign, p_zero); }
            }
        }
    }
    printf("%s\n", failures ? "FAILED" : "ok");
    return failures ? 1 : 0;
}
