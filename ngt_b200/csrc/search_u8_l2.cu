// search_u8_l2.cu -- instantiates the traversal kernels of one accumulate kind (see search.cuh).
#include "search.cuh"
template cudaError_t search_dispatch<3>(const SearchArgs &, const SearchLaunch &, int, int *);
