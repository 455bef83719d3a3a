/* TEST INFRASTRUCTURE — compiles /root/reference/bpf/dhcp_fastpath.c (+maps.h) verbatim. */
#define _license _license_dhcp
#include "dhcp_fastpath.c"
#include "ref_common.h"

REF_XDP_WRAPPER(run_dhcp_fastpath, dhcp_fastpath_prog)

const ora_map_desc ref_dhcp_maps[] = {
    REF_MAP_KV(subscriber_pools), REF_MAP_KV(vlan_subscriber_pools), REF_MAP_KV(ip_pools),
    REF_MAP_KV(server_config),    REF_MAP_KV(stats_map),             REF_MAP_KV(circuit_id_map),
    REF_MAP_KV(circuit_id_subscribers),
};
const int ref_dhcp_nmaps = sizeof(ref_dhcp_maps) / sizeof(ref_dhcp_maps[0]);
const ora_prog_desc ref_dhcp_progs[] = {{"dhcp_fastpath_prog", run_dhcp_fastpath}};
const int ref_dhcp_nprogs = 1;
