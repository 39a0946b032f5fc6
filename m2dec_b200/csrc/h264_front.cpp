// placeholder until the H.264 front end lands
#include "frontend.h"
namespace b2 {
Backend *h264_backend(void *) { return nullptr; }
void h264_destroy(void *) {}
}
extern "C" const m2d_func_table_t *const h264d_gpu_func = nullptr;
