/* TEST INFRASTRUCTURE — CPU parity oracle for the bng dataplane hot path.
 *
 * This header is shared by the two oracle builds:
 *   oracle/_ref/libbng_ref.so   the reference's own eBPF C sources
 *                               (/root/reference/bpf/{antispoof,qos_ratelimit,
 *                               nat44,dhcp_fastpath}.c) compiled natively with
 *                               gcc against oracle/shim + oracle/ref_glue.c
 *   oracle/libbng_port.so       a plain-C restatement (oracle/port.c) that is
 *                               pinned against the former
 * Both export the same ora_* C API below, on top of the same eBPF-map
 * emulation (oracle/runtime.c).  Nothing under oracle/ is product code: only
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm
 * may load these libraries.
 */
#ifndef BNG_ORACLE_API_H
#define BNG_ORACLE_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* eBPF map types that the four reference programs declare
 * (values are the kernel's enum bpf_map_type). */
enum {
    ORA_MAP_HASH = 1,
    ORA_MAP_ARRAY = 2,
    ORA_MAP_PERF_EVENT_ARRAY = 4,
    ORA_MAP_PERCPU_ARRAY = 6,
    ORA_MAP_LRU_HASH = 9,
    ORA_MAP_LPM_TRIE = 11,
    ORA_MAP_RINGBUF = 27,
};

/* bpf(2) update flags */
#define ORA_ANY 0
#define ORA_NOEXIST 1
#define ORA_EXIST 2

typedef struct ora_map_info {
    uint32_t type;
    uint32_t key_size;
    uint32_t value_size;
    uint32_t max_entries;
    uint64_t count;
} ora_map_info;

/* One batch of frames.  Frame i occupies bytes [off16[i]*16, off16[i]*16+len[i])
 * of the arena (or i*stride when off16 is NULL).  Frames are processed in
 * index order; every bpf_ktime_get_ns() issued while the batch runs returns
 * now_ns. */
typedef struct ora_batch {
    uint8_t *pkts;
    const uint32_t *off16;
    uint32_t *len;      /* in: frame length; out: length after bpf_xdp_adjust_tail */
    uint8_t *verdict;   /* out: TC_ACT_* or XDP_* code */
    uint32_t *priority; /* in/out: skb->priority (may be NULL) */
    uint32_t n;
    uint32_t stride;
    uint64_t now_ns;
    const uint64_t *now_v; /* optional: bpf_ktime_get_ns() per frame ([n]); NULL = now_ns for the whole batch */
} ora_batch;

const char *ora_impl(void); /* "reference" or "port" */
void ora_reset(void);

int ora_map_count_all(void);
const char *ora_map_name(int id);
int ora_map_id(const char *name);
int ora_map_get_info(int id, ora_map_info *out);
int ora_map_update(int id, const void *key, const void *val, uint64_t flags);
int ora_map_update_batch(int id, const void *keys, const void *vals, uint64_t n, uint64_t flags);
int ora_map_lookup(int id, const void *key, void *val_out);
int ora_map_delete(int id, const void *key);
uint64_t ora_map_dump(int id, void *keys, void *vals, uint64_t cap);

int ora_prog_id(const char *name);
const char *ora_prog_name(int id);
int ora_prog_run(int prog, ora_batch *b);

/* Records written through bpf_perf_event_output / bpf_ringbuf_submit, in
 * emission order.  Returns the number of records copied and removes them. */
uint64_t ora_events_drain(int map_id, void *buf, uint64_t cap_records);
uint32_t ora_event_size(int map_id);

/* Packet memory below 4 GiB (struct __sk_buff / xdp_md carry 32-bit data
 * pointers, /usr/include/linux/bpf.h). */
void *ora_arena_alloc(size_t bytes);
void ora_arena_free(void *p, size_t bytes);

/* ---- internal: runtime <-> implementation ---- */
typedef struct ora_map_desc {
    const char *name;
    uint32_t type, key_size, value_size, max_entries;
    void *ref_addr; /* address of the reference's map global (reference build) */
} ora_map_desc;

typedef struct ora_pkt {
    uint8_t *data;
    uint32_t len;      /* in/out */
    uint32_t priority; /* in/out */
    void *ctx_cookie;
} ora_pkt;

typedef int (*ora_prog_fn)(ora_pkt *p);

typedef struct ora_prog_desc {
    const char *name;
    ora_prog_fn fn;
} ora_prog_desc;

/* provided by ref_glue.c or port.c */
const char *ora_impl_name(void);
const ora_map_desc *ora_impl_maps(int *n);
const ora_prog_desc *ora_impl_progs(int *n);
void ora_impl_bind(void); /* called after maps are (re)created */

/* provided by runtime.c for the implementations */
struct ora_map;
struct ora_map *ora_rt_map(int id);
void *ora_rt_lookup(struct ora_map *m, const void *key);
long ora_rt_update(struct ora_map *m, const void *key, const void *val, uint64_t flags);
long ora_rt_delete(struct ora_map *m, const void *key);
uint64_t ora_rt_now(void);
long ora_rt_event_output(struct ora_map *m, const void *data, uint64_t size);
void *ora_rt_ringbuf_reserve(struct ora_map *m, uint64_t size);
void ora_rt_ringbuf_submit(void *rec);

#ifdef __cplusplus
}
#endif
#endif
