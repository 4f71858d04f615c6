// search_fast_u8_ham.cu -- instantiates the lean traversal kernels of one accumulate kind (see search_fast.cuh).
#include "search_fast.cuh"
template cudaError_t search_fast_dispatch<ACC_U8_HAM>(const SearchArgs &, int, int, unsigned, size_t, cudaStream_t, int, int *);
template cudaError_t seed_select_dispatch<ACC_U8_HAM>(const SeedArgs &, cudaStream_t);
