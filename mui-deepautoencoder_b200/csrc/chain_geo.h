// Work decomposition of the chain kernel (gemm_chain.cu), host + device: which k-blocks of which output tile a CTA owns.
// Kept free of CUDA types so that tests/test_chain_geometry.py can compile it with gcc and check coverage on the CPU.
#pragma once

#ifdef __CUDACC__
#define CHAIN_HD __host__ __device__ __forceinline__
#else
#define CHAIN_HD static inline
#endif

enum { kChainBN = 64, kChainBK = 64 };

typedef struct Geo { int tiles, total_kb, kb_begin, num_kb, nsplit; } Geo;

CHAIN_HD int chain_tiles(int N) { return (N + kChainBN - 1) / kChainBN; }

// Layer with N output columns and contraction length K, cluster of S CTAs, this CTA's rank in the cluster:
// the k-blocks are split into nsplit <= S contiguous ranges of kb_per blocks; ranks >= nsplit own none.
CHAIN_HD Geo chain_layer_geo(int N, int K, int S, int rank) {
    Geo g;
    g.tiles = chain_tiles(N);
    g.total_kb = (K + kChainBK - 1) / kChainBK;
    const int kb_per = (g.total_kb + S - 1) / S;
    g.nsplit = (g.total_kb + kb_per - 1) / kb_per;
    g.kb_begin = rank * kb_per;
    int end = g.kb_begin + kb_per;
    if (end > g.total_kb) end = g.total_kb;
    g.num_kb = end > g.kb_begin ? end - g.kb_begin : 0;
    return g;
}
