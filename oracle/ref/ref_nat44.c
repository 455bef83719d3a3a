/* TEST INFRASTRUCTURE — compiles /root/reference/bpf/nat44.c verbatim. */
#define _license _license_nat44
#include "nat44.c"
#include "ref_common.h"

REF_TC_WRAPPER(run_nat44_egress, nat44_egress)
REF_TC_WRAPPER(run_nat44_ingress, nat44_ingress)
REF_XDP_WRAPPER(run_nat44_hairpin_xdp, nat44_hairpin_xdp)

const ora_map_desc ref_nat44_maps[] = {
    REF_MAP_KV(nat_sessions),   REF_MAP_KV(nat_reverse),   REF_MAP_KV(eim_table), REF_MAP_KV(subscriber_nat),
    REF_MAP_KV(nat_pool),       REF_MAP_KV(hairpin_ips),   REF_MAP_KV(nat_config_map),
    REF_MAP_KV(nat_stats_map),  REF_MAP_RING(nat_log_rb),  REF_MAP_KV(alg_ports), REF_MAP_KV(nat_private_ranges),
};
const int ref_nat44_nmaps = sizeof(ref_nat44_maps) / sizeof(ref_nat44_maps[0]);
const ora_prog_desc ref_nat44_progs[] = {
    {"nat44_egress", run_nat44_egress},
    {"nat44_ingress", run_nat44_ingress},
    {"nat44_hairpin_xdp", run_nat44_hairpin_xdp},
};
const int ref_nat44_nprogs = 3;
