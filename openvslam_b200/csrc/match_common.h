// match_common.h -- the matcher handle shared by match_bruteforce.cu and match_window.cu.
#pragma once
#include <algorithm>
#include <vector>
#include "ovs_common.h"

// device buffer of a released frame index, kept for the next ovs_frame_index_create* (one per frame in a tracking loop:
// cudaMalloc / cudaFree per frame would serialise every stream of the device)
struct ovs_index_buf { uint8_t* base = nullptr; size_t cap = 0; };

struct ovs_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
    // grow-only device / pinned scratch
    uint8_t* d_q = nullptr; size_t d_q_cap = 0;
    uint8_t* d_t = nullptr; size_t d_t_cap = 0;
    unsigned* d_part = nullptr; size_t d_part_cap = 0;
    unsigned* d_keys = nullptr; size_t d_keys_cap = 0;
    unsigned* d_mask = nullptr; size_t d_mask_cap = 0;
    unsigned* h_keys = nullptr; size_t h_keys_cap = 0;  // pinned
    uint8_t* h_stage = nullptr; size_t h_stage_cap = 0; // pinned
    cudaEvent_t ev[2]{};
    float last_kernel_us = 0.f;
    int num_requeries = 0;   // GPU re-queries issued by the greedy replays so far (diagnostic)
    std::vector<ovs_index_buf> index_pool;
};

namespace ovs {

template <typename T>
int grow_dev(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return OVS_OK;
    cudaFree(*p); *p = nullptr; *cap = 0;
    // a quarter of slack: the number of keypoints creeps from frame to frame, and every cudaFree / cudaMalloc stalls all streams of the device
    const size_t n = std::max(need + need / 4, (size_t)4096);
    OVS_CUDA_CHECK(cudaMalloc(p, n * sizeof(T)));
    *cap = n;
    return OVS_OK;
}
template <typename T>
int grow_host(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return OVS_OK;
    cudaFreeHost(*p); *p = nullptr; *cap = 0;
    const size_t n = std::max(need + need / 4, (size_t)4096);
    OVS_CUDA_CHECK(cudaHostAlloc(p, n * sizeof(T), cudaHostAllocDefault));
    *cap = n;
    return OVS_OK;
}

}  // namespace ovs
