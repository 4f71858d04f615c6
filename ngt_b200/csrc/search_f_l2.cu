// search_f_l2.cu -- instantiates the traversal kernels of one accumulate kind (see search.cuh).
#include "search.cuh"
template cudaError_t search_dispatch<0>(const SearchArgs &, const SearchLaunch &, int, int *);
