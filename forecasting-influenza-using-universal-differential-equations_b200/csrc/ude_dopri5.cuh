// ude_dopri5.cuh — adaptive Dormand-Prince 5(4) kernels (placeholder until the RK4 path is validated).
#pragma once
#include "ude_engine.cuh"
namespace ude {
constexpr int kDopriStages = 7;
enum { DF_NSTATE = 12, DB_NSTATE = 10 };
}  // namespace ude
