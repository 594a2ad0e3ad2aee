// Host-side sampler runtime (C++): paged-KV block manager + continuous-batching scheduler.
// Replaces the scheduler / block-manager half of the vLLM engine the reference boots per rollout
// (/root/reference/GRPO/grpo_trainer.py:141-142; SURVEY.md section 2.6 first row).
#pragma once
#include <pybind11/pybind11.h>

namespace nrl {
void bind_runtime(pybind11::module_& m);
}
