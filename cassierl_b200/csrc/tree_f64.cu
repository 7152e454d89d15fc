// fp64 instantiation of the 3-D tree engine kernels (parity build)
#include "tree_kernels.cuh"
namespace cassie { namespace tree { CASSIE_TREE_INSTANTIATE(double) } }
