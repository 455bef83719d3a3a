/* TEST INFRASTRUCTURE — the 8 BPF helpers the reference programs import
 * (SURVEY.md §0), implemented over oracle/runtime.c, plus the descriptor
 * tables that tie the reference's map globals to runtime maps. */
#include <errno.h>
#include <linux/bpf.h>
#include <stdint.h>
#include <string.h>

#include "ref_common.h"

extern const ora_map_desc ref_antispoof_maps[], ref_qos_maps[], ref_nat44_maps[], ref_dhcp_maps[];
extern const int ref_antispoof_nmaps, ref_qos_nmaps, ref_nat44_nmaps, ref_dhcp_nmaps;
extern const ora_prog_desc ref_antispoof_progs[], ref_qos_progs[], ref_nat44_progs[], ref_dhcp_progs[];
extern const int ref_antispoof_nprogs, ref_qos_nprogs, ref_nat44_nprogs, ref_dhcp_nprogs;

static ora_map_desc all_maps[40];
static int n_all_maps;
static ora_prog_desc all_progs[16];
static int n_all_progs;

/* address -> runtime map, 128-slot open table */
#define AT 128
static void *at_addr[AT];
static struct ora_map *at_map[AT];

static inline unsigned at_hash(const void *a) { return (unsigned)(((uintptr_t)a >> 4) * 2654435761u) & (AT - 1); }

static inline struct ora_map *map_of(const void *addr) {
    unsigned h = at_hash(addr);
    while (at_addr[h]) {
        if (at_addr[h] == addr) return at_map[h];
        h = (h + 1) & (AT - 1);
    }
    return NULL;
}

const char *ora_impl_name(void) { return "reference"; }

static void collect(void) {
    if (n_all_maps) return;
    const ora_map_desc *ms[] = {ref_antispoof_maps, ref_qos_maps, ref_nat44_maps, ref_dhcp_maps};
    const int mn[] = {ref_antispoof_nmaps, ref_qos_nmaps, ref_nat44_nmaps, ref_dhcp_nmaps};
    const ora_prog_desc *ps[] = {ref_antispoof_progs, ref_qos_progs, ref_nat44_progs, ref_dhcp_progs};
    const int pn[] = {ref_antispoof_nprogs, ref_qos_nprogs, ref_nat44_nprogs, ref_dhcp_nprogs};
    for (int g = 0; g < 4; g++) {
        for (int i = 0; i < mn[g]; i++) all_maps[n_all_maps++] = ms[g][i];
        for (int i = 0; i < pn[g]; i++) all_progs[n_all_progs++] = ps[g][i];
    }
}

const ora_map_desc *ora_impl_maps(int *n) {
    collect();
    *n = n_all_maps;
    return all_maps;
}
const ora_prog_desc *ora_impl_progs(int *n) {
    collect();
    *n = n_all_progs;
    return all_progs;
}
void ora_impl_bind(void) {
    memset(at_addr, 0, sizeof(at_addr));
    for (int i = 0; i < n_all_maps; i++) {
        unsigned h = at_hash(all_maps[i].ref_addr);
        while (at_addr[h]) h = (h + 1) & (AT - 1);
        at_addr[h] = all_maps[i].ref_addr;
        at_map[h] = ora_rt_map(i);
    }
}

/* ---- low-memory bounce ---- */
static uint8_t *bounce;
uint8_t *ref_low_begin(ora_pkt *p) {
    if ((uintptr_t)p->data + 65536 <= 0xFFFFFFFFull) return p->data;
    if (!bounce) bounce = ora_arena_alloc(65536);
    memcpy(bounce, p->data, p->len);
    return bounce;
}
void ref_low_done(ora_pkt *p, uint8_t *low, uint32_t new_len) {
    if (low != p->data) memcpy(p->data, low, new_len > p->len ? new_len : p->len);
}

/* ---- helpers ---- */
void *bpf_map_lookup_elem(void *map, const void *key) { return ora_rt_lookup(map_of(map), key); }
long bpf_map_update_elem(void *map, const void *key, const void *value, __u64 flags) {
    return ora_rt_update(map_of(map), key, value, flags);
}
long bpf_map_delete_elem(void *map, const void *key) { return ora_rt_delete(map_of(map), key); }
__u64 bpf_ktime_get_ns(void) { return ora_rt_now(); }
long bpf_perf_event_output(void *ctx, void *map, __u64 flags, void *data, __u64 size) {
    (void)ctx;
    (void)flags;
    return ora_rt_event_output(map_of(map), data, size);
}
void *bpf_ringbuf_reserve(void *ringbuf, __u64 size, __u64 flags) {
    (void)flags;
    return ora_rt_ringbuf_reserve(map_of(ringbuf), size);
}
void bpf_ringbuf_submit(void *data, __u64 flags) {
    (void)flags;
    ora_rt_ringbuf_submit(data);
}
long bpf_xdp_adjust_tail(void *xdp, int delta) {
    struct xdp_md *x = xdp;
    int64_t nend = (int64_t)x->data_end + delta;
    if (nend < (int64_t)x->data + 14) return -EINVAL; /* net/core/filter.c: below ETH_HLEN */
    x->data_end = (uint32_t)nend;
    return 0;
}
