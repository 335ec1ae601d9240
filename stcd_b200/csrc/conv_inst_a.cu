// conv_ws_kernel instances, part A (see STCD_CONV_INSTANCES_A in conv_ws.cuh): one of six translation units
// compiled in parallel.
#include "conv_ws.cuh"

namespace stcd {
STCD_DEFINE_CONV_TABLE(conv_kernel_table_a, STCD_CONV_INSTANCES_A)
}  // namespace stcd
