// conv_ws_kernel instances, part F (see STCD_CONV_INSTANCES_F in conv_ws.cuh): the eight-epilogue-warp variants; one of six
// translation units compiled in parallel.
#include "conv_ws.cuh"

namespace stcd {
STCD_DEFINE_CONV_TABLE_WITH(conv_kernel_table_f, STCD_CONV_INSTANCES_F, STCD_CONV_TABLE_ENTRY8)
}  // namespace stcd
