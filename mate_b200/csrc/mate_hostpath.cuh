// mate_hostpath.cuh -- the device -> host leg of mate_b200_step_host.
//
// The observation rows of a step are 6.2 KB per environment (MATE-4v8-9); the plain leg copies them densely and is bound
// by the PCIe link (407 MB per step of 65 536 environments at 57 GB/s = 7.1 ms).  Two lossless alternatives, both built
// from the same pieces -- a kernel that keeps some of the rows' 16-byte chunks and records which (one bit per chunk, one
// table entry per 256 chunks), the kept chunks crossing the link as one stream per launch chunk, and a pool of host
// threads that put them where they belong in the caller's buffers while later streams are still in flight:
//   * compact_chunks_kernel keeps the chunks that are not all-zero (55 % are: entities the observer does not see); the
//     host threads rebuild every byte of the rows (zeros included);
//   * compact_changes_kernel, when the caller's buffers still hold the previous call's rows (MATE_STEP_HOST_ROWS_KEPT) and
//     the device does too, keeps the 64-byte groups that differ from them (45 % under random actions); the host threads
//     rewrite exactly those cache lines.
// The caller sees the bytes of the dense copy either way (tests/test_cuda_parity.py::test_step_host_*).  What bounds the
// leg then is the host's memory system and the number of host threads (profiles/r2v_hostpath.md; first probe:
// profiles/tools/host_expand_probe.cu).
#pragma once

#include <cuda_runtime.h>
#include <emmintrin.h>

#include <condition_variable>
#include <cstdint>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace mate {

constexpr int kCompactBlock = 256;                   // 16-byte chunks per block: one table entry
constexpr int kCompactWords = kCompactBlock / 32;    // bitmap words per block
struct CompactEntry {
    uint32_t offset;                                 // position of the block's first kept chunk in the compact stream
    uint32_t words[kCompactWords];                   // bit k of word w: chunk 32 w + k of the block is not all-zero
};

// One warp per block of 256 chunks: the block's chunks stay in registers between the vote and the packed store; the
// position in the compact stream comes from one atomicAdd per block (the order of the blocks in the stream is
// arbitrary, the table records it).
__global__ void compact_chunks_kernel(const uint4* __restrict__ src, const long long nchunks, uint4* __restrict__ dst,
                                      CompactEntry* __restrict__ table, unsigned int* __restrict__ counter) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long nblocks = (nchunks + kCompactBlock - 1) / kCompactBlock;
    for (long long b = warp; b < nblocks; b += nwarps) {
        uint4 v[kCompactWords];
        uint32_t words[kCompactWords];
        int total = 0;
#pragma unroll
        for (int it = 0; it < kCompactWords; ++it) {
            const long long k = b * kCompactBlock + it * 32 + lane;
            v[it] = k < nchunks ? src[k] : make_uint4(0u, 0u, 0u, 0u);
            words[it] = __ballot_sync(0xffffffffu, (v[it].x | v[it].y | v[it].z | v[it].w) != 0u);
            total += __popc(words[it]);
        }
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned int)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        uint32_t run = base;
#pragma unroll
        for (int it = 0; it < kCompactWords; ++it) {
            if ((words[it] >> lane) & 1u) dst[run + __popc(words[it] & ((1u << lane) - 1u))] = v[it];
            run += __popc(words[it]);
            if (lane == it) table[b].words[it] = words[it];
        }
        if (lane == 0) table[b].offset = base;
    }
}

// The same for the rows of a call against the rows of the previous call (both on the device): a 64-byte GROUP of four
// chunks is kept when any of its bytes differs, so that the host rewrites whole cache lines and nothing else.
__global__ void compact_changes_kernel(const uint4* __restrict__ src, const uint4* __restrict__ before, const long long nchunks,
                                       uint4* __restrict__ dst, CompactEntry* __restrict__ table, unsigned int* __restrict__ counter) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long nblocks = (nchunks + kCompactBlock - 1) / kCompactBlock;
    for (long long b = warp; b < nblocks; b += nwarps) {
        uint4 v[kCompactWords];
        uint32_t words[kCompactWords];
        int total = 0;
#pragma unroll
        for (int it = 0; it < kCompactWords; ++it) {
            const long long k = b * kCompactBlock + it * 32 + lane;
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            v[it] = q;
            if (k < nchunks) { v[it] = src[k]; q = before[k]; }
            uint32_t w = __ballot_sync(0xffffffffu, ((v[it].x ^ q.x) | (v[it].y ^ q.y) | (v[it].z ^ q.z) | (v[it].w ^ q.w)) != 0u);
            w |= w >> 1; w |= w >> 2;                       // bit 4 g: any chunk of group g differs
            words[it] = (w & 0x11111111u) * 0xFu;           // all four chunks of such a group
            total += __popc(words[it]);
        }
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned int)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        uint32_t run = base;
#pragma unroll
        for (int it = 0; it < kCompactWords; ++it) {
            if ((words[it] >> lane) & 1u) dst[run + __popc(words[it] & ((1u << lane) - 1u))] = v[it];
            run += __popc(words[it]);
            if (lane == it) table[b].words[it] = words[it];
        }
        if (lane == 0) table[b].offset = base;
    }
}

// the two stream sizes of a launch chunk, stored straight into pinned host memory
__global__ void publish_counts_kernel(const unsigned int* __restrict__ counts, unsigned int* __restrict__ host_counts) {
    if (threadIdx.x < 2) {
        host_counts[threadIdx.x] = counts[threadIdx.x];
        __threadfence_system();
    }
}

// Expand blocks [b0, b1) of a region: dst = the caller's dense rows (16-byte chunks), src = the compact stream.
// A group of four chunks that is all-zero or all-kept is written without looking at its chunks; in a mixed group the next
// compact chunk is always loaded and kept only if its bit is set (the stream buffer is some chunks longer than its content).
// With `only_marked` (compact_changes_kernel: whole 64-byte groups that differ from what the buffers hold) the unmarked
// groups are left alone instead of being zeroed.
template <bool ALIGNED>
inline void expand_blocks_impl(const CompactEntry* table, const bool only_marked, const __m128i* stream, __m128i* dst, long long nchunks,
                               long long b0, long long b1) {
    const __m128i zero = _mm_setzero_si128();
    auto put = [](__m128i* q, const __m128i x) { if (ALIGNED) _mm_stream_si128(q, x); else _mm_storeu_si128(q, x); };
    for (long long b = b0; b < b1; ++b) {
        const CompactEntry& entry = table[b];
        const __m128i* s = stream + entry.offset;
        __m128i* d = dst + b * kCompactBlock;
        const long long left = nchunks - b * kCompactBlock;
        for (int w = 0; w < kCompactWords; ++w) {
            const long long rem = left - 32 * w;
            if (rem <= 0) break;
            const int n = rem < 32 ? (int)rem : 32;
            const uint32_t m = entry.words[w];
            __m128i* dw = d + 32 * w;
            if (only_marked) {
                // whole groups of four chunks (the region is a whole number of groups): one pass over the marked groups
                uint32_t g = m & 0x11111111u;
                while (g) {
                    const int k0 = __builtin_ctz(g);
                    g &= g - 1u;
                    const __m128i x0 = _mm_loadu_si128(s), x1 = _mm_loadu_si128(s + 1), x2 = _mm_loadu_si128(s + 2), x3 = _mm_loadu_si128(s + 3);
                    s += 4;
                    put(dw + k0, x0); put(dw + k0 + 1, x1); put(dw + k0 + 2, x2); put(dw + k0 + 3, x3);
                }
                continue;
            }
            for (int k0 = 0; k0 < n; k0 += 4) {
                const uint32_t nib = (m >> k0) & 0xFu;
                const int k1 = k0 + 4 < n ? k0 + 4 : n;
                if (nib == 0u) {
                    for (int k = k0; k < k1; ++k) put(dw + k, zero);
                } else if (nib == 0xFu && k1 == k0 + 4) {
                    const __m128i x0 = _mm_loadu_si128(s), x1 = _mm_loadu_si128(s + 1), x2 = _mm_loadu_si128(s + 2), x3 = _mm_loadu_si128(s + 3);
                    s += 4;
                    put(dw + k0, x0); put(dw + k0 + 1, x1); put(dw + k0 + 2, x2); put(dw + k0 + 3, x3);
                } else {
                    for (int k = k0; k < k1; ++k) {   // the next compact chunk is loaded in any case and kept only if the bit is set
                        const uint32_t bit = (m >> k) & 1u;
                        const __m128i x = _mm_and_si128(_mm_loadu_si128(s), _mm_set1_epi32(-(int)bit));
                        s += bit;
                        put(dw + k, x);
                    }
                }
            }
        }
    }
    _mm_sfence();
}
inline void expand_blocks(const CompactEntry* table, const bool only_marked, const __m128i* stream, __m128i* dst, long long nchunks,
                          long long b0, long long b1, bool aligned) {
    if (aligned) expand_blocks_impl<true>(table, only_marked, stream, dst, nchunks, b0, b1);
    else expand_blocks_impl<false>(table, only_marked, stream, dst, nchunks, b0, b1);
}

// A fixed pool of host threads that expand pieces of the compact stream.
class ExpandPool {
public:
    struct Work {
        const CompactEntry* table;
        bool only_marked;                            // the table marks what differs from the buffers' content
        const __m128i* stream;
        __m128i* dst;
        long long nchunks, b0, b1;
        bool aligned;
        cudaEvent_t ready;                           // the piece's part of the stream has arrived once this event is done
    };
    ExpandPool(int threads, int device) : device_(device) {
        for (int i = 0; i < threads; ++i) pool_.emplace_back([this] { run(); });
    }
    ~ExpandPool() {
        {
            std::lock_guard<std::mutex> lock(mu_);
            quit_ = true;
        }
        cv_work_.notify_all();
        for (auto& t : pool_) t.join();
    }
    void submit(const Work& w) {
        {
            std::lock_guard<std::mutex> lock(mu_);
            queue_.push_back(w);
            ++pending_;
        }
        cv_work_.notify_one();
    }
    void wait() {
        std::unique_lock<std::mutex> lock(mu_);
        cv_done_.wait(lock, [this] { return pending_ == 0; });
    }
    int size() const { return (int)pool_.size(); }
    int pending() {
        std::lock_guard<std::mutex> lock(mu_);
        return pending_;
    }

private:
    void run() {
        cudaSetDevice(device_);
        for (;;) {
            Work w;
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_work_.wait(lock, [this] { return quit_ || !queue_.empty(); });
                if (queue_.empty()) return;
                w = queue_.front();
                queue_.pop_front();
            }
            if (w.ready) cudaEventSynchronize(w.ready);
            expand_blocks(w.table, w.only_marked, w.stream, w.dst, w.nchunks, w.b0, w.b1, w.aligned);
            {
                std::lock_guard<std::mutex> lock(mu_);
                if (--pending_ == 0) cv_done_.notify_all();
            }
        }
    }
    const int device_;
    std::vector<std::thread> pool_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    std::deque<Work> queue_;
    int pending_ = 0;
    bool quit_ = false;
};

}  // namespace mate
