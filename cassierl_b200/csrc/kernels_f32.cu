// fp32 instantiation of the batched Cassie2d kernels (the production precision).
#include "env_kernels.cuh"
namespace cassie {
template struct Launch<float>;
}
