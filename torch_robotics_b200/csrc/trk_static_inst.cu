// Explicit instantiation of the compiled-in kernel family for ONE robot (-DTRK_INST_ROBOT=robots::<Name>).
// Built once per robot listed in robots_generated.cuh, in parallel, and linked into libtrk_b200.so.
#ifndef TRK_INST_ROBOT
#error "compile with -DTRK_INST_ROBOT=robots::<Name>"
#endif
#include "trk_static.cuh"

namespace trk {
TRK_LAUNCHER_SIGNATURES(, TRK_INST_ROBOT)
}  // namespace trk
