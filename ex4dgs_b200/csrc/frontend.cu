// Fused model front-end - placeholder until the kernels land.
#include "../../include/ex4dgs_raster.h"
extern "C" {
int ex4dgs_frontend_forward(int, int, int, const float*, const float*, const float*, const float*, const float*,
                            const float*, const float*, const float*, const float*, const float*, const float*,
                            float, float, float, float, float, float*, float*, float*, float*, void*)
{
    return EX4DGS_ERR_UNSUPPORTED;
}
int ex4dgs_frontend_backward(int, int, int, const float*, const float*, const float*, const float*, const float*,
                             const float*, const float*, const float*, float, float, float, float, float,
                             const float*, const float*, const float*, const float*, float*, float*, float*, float*,
                             float*, float*, float*, float*, float*, float*, float*, void*)
{
    return EX4DGS_ERR_UNSUPPORTED;
}
}
