/* TEST INFRASTRUCTURE — glue shared by the four wrapper translation units that
 * #include the reference's eBPF C sources verbatim (found through
 * -I/root/reference/bpf; nothing is copied into this repository).
 * Map key/value sizes, types and capacities are recovered from the reference's
 * own BTF-style map definitions with sizeof(), so the oracle's tables are the
 * reference's tables by construction. */
#ifndef BNG_ORACLE_REF_COMMON_H
#define BNG_ORACLE_REF_COMMON_H
#include <stdint.h>
#include "../oracle_api.h"

#define REF_UINT(m, f) ((uint32_t)(sizeof(*(m).f) / sizeof(int)))
#define REF_MAP_KV(m) \
    { #m, REF_UINT(m, type), (uint32_t)sizeof(*(m).key), (uint32_t)sizeof(*(m).value), REF_UINT(m, max_entries), &(m) }
#define REF_MAP_PERF(m) \
    { #m, REF_UINT(m, type), REF_UINT(m, key_size), REF_UINT(m, value_size), 0, &(m) }
#define REF_MAP_RING(m) \
    { #m, REF_UINT(m, type), 0, 0, REF_UINT(m, max_entries), &(m) }

/* Returns a pointer below 4 GiB holding the frame (the frame itself, or a
 * bounce copy); ref_low_done() copies a bounce buffer back. */
uint8_t *ref_low_begin(ora_pkt *p);
void ref_low_done(ora_pkt *p, uint8_t *low, uint32_t new_len);

#define REF_TC_WRAPPER(wrapper, prog)                         \
    static int wrapper(ora_pkt *p) {                          \
        struct __sk_buff skb;                                 \
        __builtin_memset(&skb, 0, sizeof(skb));               \
        uint8_t *low = ref_low_begin(p);                      \
        skb.len = p->len;                                     \
        skb.priority = p->priority;                           \
        skb.data = (uint32_t)(uintptr_t)low;                  \
        skb.data_end = (uint32_t)(uintptr_t)low + p->len;     \
        int v = prog(&skb);                                   \
        p->priority = skb.priority;                           \
        ref_low_done(p, low, p->len);                         \
        return v;                                             \
    }

#define REF_XDP_WRAPPER(wrapper, prog)                        \
    static int wrapper(ora_pkt *p) {                          \
        struct xdp_md x;                                      \
        __builtin_memset(&x, 0, sizeof(x));                   \
        uint8_t *low = ref_low_begin(p);                      \
        x.data = (uint32_t)(uintptr_t)low;                    \
        x.data_end = (uint32_t)(uintptr_t)low + p->len;       \
        int v = prog(&x);                                     \
        uint32_t nl = x.data_end - x.data;                    \
        ref_low_done(p, low, nl);                             \
        p->len = nl;                                          \
        return v;                                             \
    }
#endif
