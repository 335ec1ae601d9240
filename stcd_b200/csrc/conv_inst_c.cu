// conv_ws_kernel instances, part C (see STCD_CONV_INSTANCES_C in conv_ws.cuh): one of six translation units
// compiled in parallel.
#include "conv_ws.cuh"

namespace stcd {
STCD_DEFINE_CONV_TABLE(conv_kernel_table_c, STCD_CONV_INSTANCES_C)
}  // namespace stcd
