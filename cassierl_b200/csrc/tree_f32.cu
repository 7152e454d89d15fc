// fp32 instantiation of the 3-D tree engine kernels
#include "tree_kernels.cuh"
namespace cassie { namespace tree { CASSIE_TREE_INSTANTIATE(float) } }
