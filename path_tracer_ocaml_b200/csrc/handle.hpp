// handle.hpp — the opaque ptb_scene handle.  Product code.
#pragma once
#include "scene.hpp"

namespace ptb {
struct DeviceState;                     // render.cu
void destroy_device_state(DeviceState *);  // render.cu
}  // namespace ptb

struct ptb_scene {
  ptb::HostScene host;
  ptb::WideBVH bvh;
  std::vector<int32_t> ref_order;  // reference list order (see ptb_scene_get_prim_order); may be empty
  bool committed = false;
  ptb::DeviceState *dev = nullptr;
  // ptb_scene_commit_multi: one replica per further device (device i+1 = replicas[i]); each shares this scene's
  // tables and tree by value and owns its device copy
  std::vector<ptb_scene *> replicas;
};
