// Count-table export: compaction of the open-addressing table into an array of trew_entry, device-side sort by
// (table, k, seq) and the cross-rank merge kernel.  Stands in for the end of buffer_task* (the six ResultMaps
// handed to process_output, src/kmer.cpp:1486-1515).
#include "scan_kernels.cuh"

#include <cub/device/device_merge_sort.cuh>

namespace trew {

typedef unsigned int u32;
typedef unsigned long long u64;

// slot -> entry; unused slots are skipped, the survivors are appended in arbitrary order
__global__ void compact_kernel(const Slot* __restrict__ slots, u32 n_slots, trew_entry* __restrict__ out, u32* __restrict__ d_n, u32 cap) {
    const u32 lane = threadIdx.x & 31;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) {
        Slot s = slots[i];
        bool used = s.state == 2u && s.count != 0;
        u32 m = __ballot_sync(0xffffffffu, used);
        if (m) {
            u32 base = 0;
            if (lane == (u32)(__ffs(m) - 1)) base = atomicAdd(d_n, (u32)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            u32 o = base + __popc(m & ((1u << lane) - 1u));
            if (used && o < cap) {  // past the end of the array: count only, the host grows it and runs again
                trew_entry e;
                e.seq_lo = s.seq_lo; e.seq_hi = s.seq_hi; e.count = s.count;
                e.table = (int)(s.meta >> 8); e.k = (int)(s.meta & 0xffu);
                out[o] = e;
            }
        }
    }
}

struct EntryLess {
    __host__ __device__ bool operator()(const trew_entry& a, const trew_entry& b) const {
        if (a.table != b.table) return a.table < b.table;
        if (a.k != b.k) return a.k < b.k;
        if (a.seq_hi != b.seq_hi) return a.seq_hi < b.seq_hi;
        return a.seq_lo < b.seq_lo;
    }
};

void launch_compact(const Slot* slots, unsigned int n_slots, trew_entry* out, unsigned int* d_n, cudaStream_t stream, unsigned int cap) {
    // n_slots is a power of two >= 1024, so every warp iterates the same number of times
    compact_kernel<<<592, 256, 0, stream>>>(slots, n_slots, out, d_n, cap);
}

cudaError_t sort_entries(trew_entry* d_entries, unsigned int n, void* d_temp, size_t* temp_bytes, cudaStream_t stream) {
    return cub::DeviceMergeSort::SortKeys(d_temp, *temp_bytes, d_entries, (int)n, EntryLess(), stream);
}

}  // namespace trew
