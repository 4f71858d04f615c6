// search_f_cos.cu -- instantiates the traversal kernels of one accumulate kind (see search.cuh).
#include "search.cuh"
template cudaError_t search_dispatch<2>(const SearchArgs &, const SearchLaunch &, int, int *);
