/* TEST INFRASTRUCTURE — compiles /root/reference/bpf/antispoof.c verbatim. */
#define _license _license_antispoof
#include "antispoof.c"
#include "ref_common.h"

REF_TC_WRAPPER(run_antispoof_ingress, antispoof_ingress)

const ora_map_desc ref_antispoof_maps[] = {
    REF_MAP_KV(subscriber_bindings), REF_MAP_KV(antispoof_config), REF_MAP_KV(antispoof_stats),
    REF_MAP_PERF(spoof_events),      REF_MAP_KV(allowed_ranges_v4),
};
const int ref_antispoof_nmaps = sizeof(ref_antispoof_maps) / sizeof(ref_antispoof_maps[0]);
const ora_prog_desc ref_antispoof_progs[] = {{"antispoof_ingress", run_antispoof_ingress}};
const int ref_antispoof_nprogs = 1;
