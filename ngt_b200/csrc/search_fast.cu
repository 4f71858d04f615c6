// search_fast.cu -- instantiates the lean traversal kernels (see search_fast.cuh), one per accumulate kind and row width.
#include "search_fast.cuh"
template cudaError_t search_fast_dispatch<ACC_F_L2>(const SearchArgs &, int, unsigned, size_t, cudaStream_t, int, int *);
template cudaError_t search_fast_dispatch<ACC_F_DOT>(const SearchArgs &, int, unsigned, size_t, cudaStream_t, int, int *);
template cudaError_t search_fast_dispatch<ACC_F_COS>(const SearchArgs &, int, unsigned, size_t, cudaStream_t, int, int *);
template cudaError_t search_fast_dispatch<ACC_U8_L2>(const SearchArgs &, int, unsigned, size_t, cudaStream_t, int, int *);
template cudaError_t search_fast_dispatch<ACC_U8_HAM>(const SearchArgs &, int, unsigned, size_t, cudaStream_t, int, int *);
template cudaError_t seed_select_dispatch<ACC_F_L2>(const SeedArgs &, cudaStream_t);
template cudaError_t seed_select_dispatch<ACC_F_DOT>(const SeedArgs &, cudaStream_t);
template cudaError_t seed_select_dispatch<ACC_F_COS>(const SeedArgs &, cudaStream_t);
template cudaError_t seed_select_dispatch<ACC_U8_L2>(const SeedArgs &, cudaStream_t);
template cudaError_t seed_select_dispatch<ACC_U8_HAM>(const SeedArgs &, cudaStream_t);
