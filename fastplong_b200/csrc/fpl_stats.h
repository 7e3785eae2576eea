// Workspace of launch_cycle_stats (the length-sorted segment list), owned by a context: one stream, one workspace.
#pragma once
#include <cstddef>
#include <cstdint>

struct CycleWs {
    void* tmp = nullptr; size_t tmp_bytes = 0;
    uint32_t *k_in = nullptr, *k_out = nullptr;
    int32_t *v_in = nullptr, *v_out = nullptr;
    void* sorted = nullptr;
    int64_t cap = 0;
};
void fpl_cycle_ws_free(CycleWs* ws);
