// fp32 instantiation of the batched Cassie2d kernels (the production precision).
#include "launch.cuh"
namespace cassie {
template struct Launch<float>;
}
