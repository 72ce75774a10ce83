// Host-side planner: turns a cpz_model_desc into the device layout (ModelD) for a given column tile / block size.
#pragma once
#include <string>

#include "../../include/cpz.h"
#include "cpz_model.h"

namespace cpz {

struct PlanOptions {
  int CT = 32;            // columns per CTA
  int NT = 256;           // threads per CTA
  bool keep_all = false;  // adjoint plan: every layer output keeps its own arena rows (no aliasing), TO in {4,8}
  size_t smem_budget = 227 * 1024;
  size_t other_smem_bytes = 0;  // shared memory the kernel needs besides weights + activation arena
};

struct Plan {
  ModelD M;
  size_t smem_weight_bytes = 0;
  size_t arena_bytes = 0;  // activation arena (x2 when the kernel also keeps pre-activations)
  bool layer_major = true;
};

// Returns false and sets err on unsupported configurations.
bool validate_desc(const cpz_model_desc& d, std::string& err);
bool build_plan(const cpz_model_desc& d, const PlanOptions& opt, Plan& out, std::string& err);
void fill_tableau(int integrator, TableauD& t);
size_t count_params(const cpz_model_desc& d);

}  // namespace cpz
