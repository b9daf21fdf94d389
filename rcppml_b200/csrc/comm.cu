// comm.cu — multi-GPU (column-sharded) ALS iteration. Filled in after the single-GPU path.
#include "engine.hpp"

namespace b200 {

void Engine::comm_init(int, int, const char*) { throw std::runtime_error("multi-GPU path not built yet"); }
void Engine::comm_destroy() {}
void Engine::enqueue_iteration_sharded() { throw std::runtime_error("multi-GPU path not built yet"); }

}  // namespace b200

extern "C" {
int rcppml_b200_nccl_unique_id(char*) { b200::g_last_error = "multi-GPU path not built yet"; return -1; }
int rcppml_b200_comm_init(rcppml_b200_engine*, int, int, const char*) { b200::g_last_error = "multi-GPU path not built yet"; return -1; }
}
