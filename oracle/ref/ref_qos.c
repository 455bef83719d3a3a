/* TEST INFRASTRUCTURE — compiles /root/reference/bpf/qos_ratelimit.c verbatim. */
#define _license _license_qos
#include "qos_ratelimit.c"
#include "ref_common.h"

REF_TC_WRAPPER(run_qos_egress, qos_egress_prog)
REF_TC_WRAPPER(run_qos_ingress, qos_ingress_prog)

const ora_map_desc ref_qos_maps[] = {
    REF_MAP_KV(qos_egress), REF_MAP_KV(qos_ingress), REF_MAP_KV(qos_stats_map),
};
const int ref_qos_nmaps = sizeof(ref_qos_maps) / sizeof(ref_qos_maps[0]);
const ora_prog_desc ref_qos_progs[] = {{"qos_egress_prog", run_qos_egress}, {"qos_ingress_prog", run_qos_ingress}};
const int ref_qos_nprogs = 2;
