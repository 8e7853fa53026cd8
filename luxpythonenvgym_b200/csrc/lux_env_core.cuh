// lux_env_core.cuh -- AgentPolicy action codes -> packed action words for ONE entity
// (examples/agent_policy.py:117-132 tables, :24-100 smart_transfer_to_nearby).  Host-compilable under
// LUX_EMULATE (tests/emu) like the other *_core.cuh files.
#pragma once
#include <stdint.h>

#include "../../include/lux_b200.h"
#include "lux_warp.cuh"

#define ENV_CAP_WORKER 100
#define ENV_CAP_CART 2000

LUX_DEV int env_space(const LuxUnit& u) {
    return ((u.team_type & 2) ? ENV_CAP_CART : ENV_CAP_WORKER) - ((int)u.wood + (int)u.coal + (int)u.uranium);
}

// smart_transfer_to_nearby: largest cargo type (dict order wood, uranium, coal; strict >), best unit of
// the wanted type on the N, E, S, W cells (units of one cell in slot order)
LUX_DEV uint32_t env_smart_transfer(const LuxUnit* units, int n_units, int S, int i, int want_cart) {
    const LuxUnit U = units[i];
    const int team = U.team_type & 1;
    const int amounts[3] = {U.wood, U.uranium, U.coal};
    const int codes[3] = {0, 2, 1};
    int rtype = -1, ramount = 0;
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (amounts[k] > ramount) { rtype = codes[k]; ramount = amounts[k]; }
    int target = -1, st = 0;
    for (int k = 0; k < 4; k++) {
        int cx = U.x + (k == 1 ? 1 : k == 3 ? -1 : 0), cy = U.y + (k == 0 ? -1 : k == 2 ? 1 : 0);
        if (cx < 0 || cy < 0 || cx >= S || cy >= S) continue;
        for (int j = 0; j < n_units; j++) {
            const LuxUnit V = units[j];
            if (V.x != cx || V.y != cy || ((V.team_type >> 1) & 1) != want_cart || (V.team_type & 1) != team) continue;
            int sv = env_space(V);
            if (target < 0) { target = j; st = sv; continue; }
            if (sv >= ramount && st >= ramount) { if (sv < st) { target = j; st = sv; } }
            else if (st >= ramount) { }
            else if (sv > st) { target = j; st = sv; }
        }
    }
    if (target < 0 || rtype < 0) return 0;
    if (st < ramount) ramount = st;
    if (ramount > 2047) ramount = 2047;
    return LUX_UA_TRANSFER | ((uint32_t)rtype << 3) | ((uint32_t)ramount << 6) | ((uint32_t)units[target].id << 17);
}

// one unit decision: code index into AgentPolicy.actions_units (:117-127)
LUX_DEV uint32_t env_unit_word(const LuxUnit* units, int n_units, int S, int slot, int code) {
    const int c = ((code % 9) + 9) % 9;
    if (c == 0) return LUX_UA_MOVE | (LUX_DIR_CENTER << 3);
    if (c == 1) return LUX_UA_MOVE | (LUX_DIR_NORTH << 3);
    if (c == 2) return LUX_UA_MOVE | (LUX_DIR_WEST << 3);
    if (c == 3) return LUX_UA_MOVE | (LUX_DIR_SOUTH << 3);
    if (c == 4) return LUX_UA_MOVE | (LUX_DIR_EAST << 3);
    if (c == 5) return env_smart_transfer(units, n_units, S, slot, 1);
    if (c == 6) return env_smart_transfer(units, n_units, S, slot, 0);
    if (c == 7) return LUX_UA_BCITY;
    return LUX_UA_PILLAGE;
}
