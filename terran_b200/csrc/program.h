// Layer program (buffers, ops, weight blob) as built by program.cu from a reference state_dict.
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/terran_b200.h"

namespace trb {

struct Program {
  std::vector<tr_buffer_desc> buffers;
  std::vector<tr_op_desc> ops;
  std::vector<uint8_t> blob;
  // retinaface: head buffers of stride 32, 16, 8; arcface: embedding buffer;
  // openpose: maps buffer, PAF channel offset, heat-map channel offset
  int roles[8];
};

// model: "retinaface" | "arcface" | "openpose"; flags bit 0 (retinaface): unfused layer program
void program_build(const char* model, const void* state_dict_blob, size_t bytes, int flags, Program& out);

}  // namespace trb
