// search_fast_f_l2.cu -- instantiates the lean traversal kernels of one accumulate kind (see search_fast.cuh).
#include "search_fast.cuh"
template cudaError_t search_fast_dispatch<ACC_F_L2>(const SearchArgs &, int, int, unsigned, size_t, cudaStream_t, int, int *);
template cudaError_t seed_select_dispatch<ACC_F_L2>(const SeedArgs &, cudaStream_t);
