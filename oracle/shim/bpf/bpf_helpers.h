/* TEST INFRASTRUCTURE — userspace stand-in for libbpf's <bpf/bpf_helpers.h>.
 *
 * Lets the reference's eBPF C sources (the .c files under /root/reference/bpf) compile
 * unmodified with gcc so they can run natively as the parity oracle
 * (oracle/_ref).  Only the macros and the 8 helper prototypes those sources
 * use are provided; the helpers themselves live in oracle/runtime.c +
 * oracle/ref_glue.c.
 */
#ifndef BNG_ORACLE_SHIM_BPF_HELPERS_H
#define BNG_ORACLE_SHIM_BPF_HELPERS_H

#include <linux/types.h>
#include <stddef.h>

#define SEC(name) __attribute__((section(name), used))
#ifndef __always_inline
#define __always_inline inline __attribute__((always_inline))
#endif

/* BTF-style map definitions: sizes are recoverable with sizeof(). */
#define __uint(name, val) int(*name)[val]
#define __type(name, val) typeof(val) *name

void *bpf_map_lookup_elem(void *map, const void *key);
long bpf_map_update_elem(void *map, const void *key, const void *value, __u64 flags);
long bpf_map_delete_elem(void *map, const void *key);
__u64 bpf_ktime_get_ns(void);
long bpf_perf_event_output(void *ctx, void *map, __u64 flags, void *data, __u64 size);
void *bpf_ringbuf_reserve(void *ringbuf, __u64 size, __u64 flags);
void bpf_ringbuf_submit(void *data, __u64 flags);
long bpf_xdp_adjust_tail(void *xdp_md, int delta);

#endif
