// search_u8_ham.cu -- instantiates the traversal kernels of one accumulate kind (see search.cuh).
#include "search.cuh"
template cudaError_t search_dispatch<4>(const SearchArgs &, const SearchLaunch &, int, int *);
