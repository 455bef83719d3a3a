/* TEST INFRASTRUCTURE — userspace emulation of the eBPF map/helper runtime the
 * reference's four dataplane programs run against, plus a batch runner.
 *
 * Shared by oracle/_ref (reference C sources) and oracle/port.c (restatement).
 * Only lookup-visible semantics are modelled: HASH/LRU_HASH as
 * a node-based chained hash (value pointers stay valid across later inserts,
 * which bpf/nat44.c:516-518,730-743 relies on), ARRAY/PERCPU_ARRAY as a flat
 * single-CPU array, LPM_TRIE as longest-prefix over key bytes in memory order,
 * RINGBUF with the kernel's "8-byte header, 8-byte rounding, fail when full"
 * accounting, PERF_EVENT_ARRAY as an unbounded record log.  LRU eviction is
 * NOT modelled: an insert into a full LRU map fails with -E2BIG and bumps
 * ora_lru_overflow (tests assert it stays 0).
 */
#define _GNU_SOURCE
#include "oracle_api.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#define MAX_MAPS 40
#define MAX_PROGS 16

struct node {
    struct node *next;
    uint64_t hash;
    uint8_t kv[];
};

struct lpm_ent {
    uint32_t prefixlen;
    uint8_t *data; /* key_size-4 bytes followed by value */
};

struct ora_map {
    ora_map_desc d;
    int id;
    uint32_t voff; /* value offset inside node->kv */
    /* hash */
    struct node **buckets;
    uint64_t nb;
    uint64_t count;
    /* array */
    uint8_t *arr;
    /* lpm */
    struct lpm_ent *lpm;
    uint32_t lpm_n, lpm_cap;
    /* events (perf array / ringbuf) */
    uint8_t *ev;
    uint64_t ev_len, ev_cap, ev_nrec;
    uint32_t ev_rec_size;
    uint64_t ring_used;
    uint8_t *pending;
    uint64_t pending_size;
};

static struct ora_map g_maps[MAX_MAPS];
static int g_nmaps;
static const ora_prog_desc *g_progs;
static int g_nprogs;
static uint64_t g_now;
static int g_inited;
uint64_t ora_lru_overflow;

static int pipe_up_id = -1, pipe_tc_id = -1;
static ora_prog_fn fn_antispoof, fn_nat_egress, fn_qos_ingress;
static uint8_t *g_scratch; /* low-memory copy of the pre-NAT header for the pipeline */

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

static uint64_t hash_key(const uint8_t *k, uint32_t n) {
    uint64_t h = 0x9e3779b97f4a7c15ULL ^ n;
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, k, 8);
        h = mix64(h ^ w);
        k += 8;
        n -= 8;
    }
    if (n) {
        uint64_t w = 0;
        memcpy(&w, k, n);
        h = mix64(h ^ w ^ ((uint64_t)n << 56));
    }
    return h;
}

static int is_hash(const struct ora_map *m) {
    return m->d.type == ORA_MAP_HASH || m->d.type == ORA_MAP_LRU_HASH;
}
static int is_array(const struct ora_map *m) {
    return m->d.type == ORA_MAP_ARRAY || m->d.type == ORA_MAP_PERCPU_ARRAY;
}
static int is_event(const struct ora_map *m) {
    return m->d.type == ORA_MAP_RINGBUF || m->d.type == ORA_MAP_PERF_EVENT_ARRAY;
}

static void map_free(struct ora_map *m) {
    if (m->buckets) {
        for (uint64_t b = 0; b < m->nb; b++) {
            struct node *n = m->buckets[b];
            while (n) {
                struct node *nx = n->next;
                free(n);
                n = nx;
            }
        }
        free(m->buckets);
    }
    free(m->arr);
    for (uint32_t i = 0; i < m->lpm_n; i++) free(m->lpm[i].data);
    free(m->lpm);
    free(m->ev);
    free(m->pending);
    ora_map_desc d = m->d;
    int id = m->id;
    memset(m, 0, sizeof(*m));
    m->d = d;
    m->id = id;
}

static void map_init(struct ora_map *m) {
    m->voff = (m->d.key_size + 7u) & ~7u;
    if (is_hash(m)) {
        m->nb = 1024;
        m->buckets = calloc(m->nb, sizeof(*m->buckets));
    } else if (is_array(m)) {
        m->arr = calloc((size_t)m->d.max_entries, m->d.value_size ? m->d.value_size : 1);
    }
}

static void ensure_init(void) {
    if (g_inited) return;
    g_inited = 1;
    int n = 0;
    const ora_map_desc *d = ora_impl_maps(&n);
    if (n > MAX_MAPS) abort();
    g_nmaps = n;
    for (int i = 0; i < n; i++) {
        g_maps[i].d = d[i];
        g_maps[i].id = i;
        map_init(&g_maps[i]);
    }
    g_progs = ora_impl_progs(&g_nprogs);
    for (int i = 0; i < g_nprogs; i++) {
        if (!strcmp(g_progs[i].name, "antispoof_ingress")) fn_antispoof = g_progs[i].fn;
        if (!strcmp(g_progs[i].name, "nat44_egress")) fn_nat_egress = g_progs[i].fn;
        if (!strcmp(g_progs[i].name, "qos_ingress_prog")) fn_qos_ingress = g_progs[i].fn;
    }
    pipe_up_id = g_nprogs;
    pipe_tc_id = g_nprogs + 1;
    g_scratch = ora_arena_alloc(4096);
    ora_impl_bind();
}

const char *ora_impl(void) { return ora_impl_name(); }

void ora_reset(void) {
    ensure_init();
    for (int i = 0; i < g_nmaps; i++) {
        map_free(&g_maps[i]);
        map_init(&g_maps[i]);
    }
    g_now = 0;
    ora_lru_overflow = 0;
    ora_impl_bind();
}

int ora_map_count_all(void) {
    ensure_init();
    return g_nmaps;
}
const char *ora_map_name(int id) {
    ensure_init();
    return (id >= 0 && id < g_nmaps) ? g_maps[id].d.name : NULL;
}
int ora_map_id(const char *name) {
    ensure_init();
    for (int i = 0; i < g_nmaps; i++)
        if (!strcmp(g_maps[i].d.name, name)) return i;
    return -1;
}
struct ora_map *ora_rt_map(int id) {
    ensure_init();
    return (id >= 0 && id < g_nmaps) ? &g_maps[id] : NULL;
}
int ora_map_get_info(int id, ora_map_info *out) {
    struct ora_map *m = ora_rt_map(id);
    if (!m) return -EINVAL;
    out->type = m->d.type;
    out->key_size = m->d.key_size;
    out->value_size = m->d.value_size;
    out->max_entries = m->d.max_entries;
    out->count = is_hash(m) ? m->count : m->d.type == ORA_MAP_LPM_TRIE ? m->lpm_n
                 : is_event(m)                                         ? m->ev_nrec
                                                                       : m->d.max_entries;
    return 0;
}

/* ---------------- hash ---------------- */
static void hash_grow(struct ora_map *m) {
    uint64_t nnb = m->nb * 4;
    struct node **nb = calloc(nnb, sizeof(*nb));
    for (uint64_t b = 0; b < m->nb; b++) {
        struct node *n = m->buckets[b];
        while (n) {
            struct node *nx = n->next;
            uint64_t i = n->hash & (nnb - 1);
            n->next = nb[i];
            nb[i] = n;
            n = nx;
        }
    }
    free(m->buckets);
    m->buckets = nb;
    m->nb = nnb;
}

static struct node *hash_find(struct ora_map *m, const void *key, uint64_t h) {
    struct node *n = m->buckets[h & (m->nb - 1)];
    uint32_t ks = m->d.key_size;
    while (n) {
        if (n->hash == h && !memcmp(n->kv, key, ks)) return n;
        n = n->next;
    }
    return NULL;
}

/* ---------------- lpm ---------------- */
static int prefix_match(const uint8_t *a, const uint8_t *b, uint32_t bits) {
    uint32_t full = bits / 8, rem = bits % 8;
    if (full && memcmp(a, b, full)) return 0;
    if (rem) {
        uint8_t mask = (uint8_t)(0xff << (8 - rem));
        if ((a[full] ^ b[full]) & mask) return 0;
    }
    return 1;
}

static struct lpm_ent *lpm_exact(struct ora_map *m, const void *key) {
    uint32_t pl;
    memcpy(&pl, key, 4);
    const uint8_t *data = (const uint8_t *)key + 4;
    for (uint32_t i = 0; i < m->lpm_n; i++)
        if (m->lpm[i].prefixlen == pl && prefix_match(m->lpm[i].data, data, pl)) return &m->lpm[i];
    return NULL;
}

static void *lpm_lookup(struct ora_map *m, const void *key) {
    uint32_t pl;
    memcpy(&pl, key, 4);
    const uint8_t *data = (const uint8_t *)key + 4;
    uint32_t dbytes = m->d.key_size - 4;
    if (pl > dbytes * 8) pl = dbytes * 8;
    struct lpm_ent *best = NULL;
    for (uint32_t i = 0; i < m->lpm_n; i++) {
        struct lpm_ent *e = &m->lpm[i];
        if (e->prefixlen > pl) continue;
        if (best && e->prefixlen <= best->prefixlen) continue;
        if (prefix_match(e->data, data, e->prefixlen)) best = e;
    }
    return best ? best->data + dbytes : NULL;
}

/* ---------------- generic ops ---------------- */
void *ora_rt_lookup(struct ora_map *m, const void *key) {
    if (is_hash(m)) {
        struct node *n = hash_find(m, key, hash_key(key, m->d.key_size));
        return n ? n->kv + m->voff : NULL;
    }
    if (is_array(m)) {
        uint32_t idx;
        memcpy(&idx, key, 4);
        if (idx >= m->d.max_entries) return NULL;
        return m->arr + (size_t)idx * m->d.value_size;
    }
    if (m->d.type == ORA_MAP_LPM_TRIE) return lpm_lookup(m, key);
    return NULL;
}

long ora_rt_update(struct ora_map *m, const void *key, const void *val, uint64_t flags) {
    if (flags > ORA_EXIST) return -EINVAL;
    if (is_hash(m)) {
        uint64_t h = hash_key(key, m->d.key_size);
        struct node *n = hash_find(m, key, h);
        if (n) {
            if (flags == ORA_NOEXIST) return -EEXIST;
            memcpy(n->kv + m->voff, val, m->d.value_size);
            return 0;
        }
        if (flags == ORA_EXIST) return -ENOENT;
        if (m->count >= m->d.max_entries) {
            if (m->d.type == ORA_MAP_LRU_HASH) ora_lru_overflow++;
            return -E2BIG;
        }
        if (m->count >= m->nb * 2) hash_grow(m);
        n = malloc(sizeof(*n) + m->voff + ((m->d.value_size + 7u) & ~7u));
        n->hash = h;
        memset(n->kv, 0, m->voff);
        memcpy(n->kv, key, m->d.key_size);
        memcpy(n->kv + m->voff, val, m->d.value_size);
        uint64_t b = h & (m->nb - 1);
        n->next = m->buckets[b];
        m->buckets[b] = n;
        m->count++;
        return 0;
    }
    if (is_array(m)) {
        uint32_t idx;
        memcpy(&idx, key, 4);
        if (idx >= m->d.max_entries) return -E2BIG;
        if (flags == ORA_NOEXIST) return -EEXIST;
        memcpy(m->arr + (size_t)idx * m->d.value_size, val, m->d.value_size);
        return 0;
    }
    if (m->d.type == ORA_MAP_LPM_TRIE) {
        uint32_t pl;
        memcpy(&pl, key, 4);
        uint32_t dbytes = m->d.key_size - 4;
        if (pl > dbytes * 8) return -EINVAL;
        struct lpm_ent *e = lpm_exact(m, key);
        if (e) {
            if (flags == ORA_NOEXIST) return -EEXIST;
            memcpy(e->data + dbytes, val, m->d.value_size);
            return 0;
        }
        if (flags == ORA_EXIST) return -ENOENT;
        if (m->lpm_n >= m->d.max_entries) return -ENOSPC;
        if (m->lpm_n == m->lpm_cap) {
            m->lpm_cap = m->lpm_cap ? m->lpm_cap * 2 : 16;
            m->lpm = realloc(m->lpm, m->lpm_cap * sizeof(*m->lpm));
        }
        e = &m->lpm[m->lpm_n++];
        e->prefixlen = pl;
        e->data = malloc(dbytes + m->d.value_size);
        memcpy(e->data, (const uint8_t *)key + 4, dbytes);
        memcpy(e->data + dbytes, val, m->d.value_size);
        return 0;
    }
    return -EINVAL;
}

long ora_rt_delete(struct ora_map *m, const void *key) {
    if (is_hash(m)) {
        uint64_t h = hash_key(key, m->d.key_size);
        struct node **pp = &m->buckets[h & (m->nb - 1)];
        while (*pp) {
            struct node *n = *pp;
            if (n->hash == h && !memcmp(n->kv, key, m->d.key_size)) {
                *pp = n->next;
                free(n);
                m->count--;
                return 0;
            }
            pp = &n->next;
        }
        return -ENOENT;
    }
    if (m->d.type == ORA_MAP_LPM_TRIE) {
        struct lpm_ent *e = lpm_exact(m, key);
        if (!e) return -ENOENT;
        free(e->data);
        *e = m->lpm[--m->lpm_n];
        return 0;
    }
    return -EINVAL;
}

int ora_map_update(int id, const void *key, const void *val, uint64_t flags) {
    struct ora_map *m = ora_rt_map(id);
    if (!m || is_event(m)) return -EINVAL;
    return (int)ora_rt_update(m, key, val, flags);
}
int ora_map_update_batch(int id, const void *keys, const void *vals, uint64_t n, uint64_t flags) {
    struct ora_map *m = ora_rt_map(id);
    if (!m || is_event(m)) return -EINVAL;
    for (uint64_t i = 0; i < n; i++) {
        long r = ora_rt_update(m, (const uint8_t *)keys + i * m->d.key_size,
                               (const uint8_t *)vals + i * m->d.value_size, flags);
        if (r) return (int)r;
    }
    return 0;
}
int ora_map_lookup(int id, const void *key, void *val_out) {
    struct ora_map *m = ora_rt_map(id);
    if (!m || is_event(m)) return -EINVAL;
    void *v = ora_rt_lookup(m, key);
    if (!v) return -ENOENT;
    memcpy(val_out, v, m->d.value_size);
    return 0;
}
int ora_map_delete(int id, const void *key) {
    struct ora_map *m = ora_rt_map(id);
    if (!m || is_event(m)) return -EINVAL;
    return (int)ora_rt_delete(m, key);
}
uint64_t ora_map_dump(int id, void *keys, void *vals, uint64_t cap) {
    struct ora_map *m = ora_rt_map(id);
    if (!m) return 0;
    uint64_t k = 0;
    uint32_t ks = m->d.key_size, vs = m->d.value_size;
    if (is_hash(m)) {
        for (uint64_t b = 0; b < m->nb; b++)
            for (struct node *n = m->buckets[b]; n; n = n->next) {
                if (k >= cap) return k;
                memcpy((uint8_t *)keys + k * ks, n->kv, ks);
                memcpy((uint8_t *)vals + k * vs, n->kv + m->voff, vs);
                k++;
            }
    } else if (is_array(m)) {
        for (uint32_t i = 0; i < m->d.max_entries && k < cap; i++, k++) {
            memcpy((uint8_t *)keys + k * 4, &i, 4);
            memcpy((uint8_t *)vals + k * vs, m->arr + (size_t)i * vs, vs);
        }
    } else if (m->d.type == ORA_MAP_LPM_TRIE) {
        for (uint32_t i = 0; i < m->lpm_n && k < cap; i++, k++) {
            memcpy((uint8_t *)keys + k * ks, &m->lpm[i].prefixlen, 4);
            memcpy((uint8_t *)keys + k * ks + 4, m->lpm[i].data, ks - 4);
            memcpy((uint8_t *)vals + k * vs, m->lpm[i].data + ks - 4, vs);
        }
    }
    return k;
}

/* ---------------- clock & events ---------------- */
uint64_t ora_rt_now(void) { return g_now; }

static uint8_t *ev_reserve(struct ora_map *m, uint64_t size) {
    if (!m->ev_rec_size) m->ev_rec_size = (uint32_t)size;
    if (m->ev_len + size > m->ev_cap) {
        m->ev_cap = m->ev_cap ? m->ev_cap * 2 : 4096;
        while (m->ev_cap < m->ev_len + size) m->ev_cap *= 2;
        m->ev = realloc(m->ev, m->ev_cap);
    }
    return m->ev + m->ev_len;
}

long ora_rt_event_output(struct ora_map *m, const void *data, uint64_t size) {
    uint8_t *p = ev_reserve(m, size);
    memcpy(p, data, size);
    m->ev_len += size;
    m->ev_nrec++;
    return 0;
}

/* kernel/bpf/ringbuf.c accounting (not part of the reference tree): each
 * record carries an 8-byte header and is rounded up to 8 bytes; a reserve
 * fails when producer-consumer distance would exceed size-1. */
void *ora_rt_ringbuf_reserve(struct ora_map *m, uint64_t size) {
    uint64_t need = (size + 8 + 7) & ~7ULL;
    if (m->ring_used + need > (uint64_t)m->d.max_entries - 1) return NULL;
    if (m->pending_size < size) {
        m->pending = realloc(m->pending, size);
        m->pending_size = size;
    }
    m->ring_used += need;
    m->ev_rec_size = m->ev_rec_size ? m->ev_rec_size : (uint32_t)size;
    /* tag the pending record with its owner so submit can find the map */
    return m->pending;
}

void ora_rt_ringbuf_submit(void *rec) {
    for (int i = 0; i < g_nmaps; i++) {
        struct ora_map *m = &g_maps[i];
        if (m->d.type == ORA_MAP_RINGBUF && m->pending == rec) {
            uint8_t *p = ev_reserve(m, m->ev_rec_size);
            memcpy(p, rec, m->ev_rec_size);
            m->ev_len += m->ev_rec_size;
            m->ev_nrec++;
            return;
        }
    }
}

uint32_t ora_event_size(int map_id) {
    struct ora_map *m = ora_rt_map(map_id);
    return m ? m->ev_rec_size : 0;
}

uint64_t ora_events_drain(int map_id, void *buf, uint64_t cap_records) {
    struct ora_map *m = ora_rt_map(map_id);
    if (!m || !is_event(m) || !m->ev_nrec) return 0;
    uint64_t n = m->ev_nrec < cap_records ? m->ev_nrec : cap_records;
    uint64_t bytes = n * m->ev_rec_size;
    memcpy(buf, m->ev, bytes);
    memmove(m->ev, m->ev + bytes, m->ev_len - bytes);
    m->ev_len -= bytes;
    m->ev_nrec -= n;
    if (m->d.type == ORA_MAP_RINGBUF) {
        uint64_t per = ((uint64_t)m->ev_rec_size + 8 + 7) & ~7ULL;
        m->ring_used -= n * per;
    }
    return n;
}

/* ---------------- arena ---------------- */
void *ora_arena_alloc(size_t bytes) {
    void *p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_32BIT, -1, 0);
    if (p == MAP_FAILED) return NULL;
    return p;
}
void ora_arena_free(void *p, size_t bytes) {
    if (p) munmap(p, bytes);
}

/* ---------------- programs ---------------- */
int ora_prog_id(const char *name) {
    ensure_init();
    for (int i = 0; i < g_nprogs; i++)
        if (!strcmp(g_progs[i].name, name)) return i;
    if (!strcmp(name, "pipeline_up")) return pipe_up_id;
    if (!strcmp(name, "pipeline_tc")) return pipe_tc_id;
    return -1;
}
const char *ora_prog_name(int id) {
    ensure_init();
    if (id >= 0 && id < g_nprogs) return g_progs[id].name;
    if (id == pipe_up_id) return "pipeline_up";
    if (id == pipe_tc_id) return "pipeline_tc";
    return NULL;
}

/* Upstream pipeline compositions.  TC_ACT_SHOT (2) at any
 * stage ends processing of that frame.
 *   pipeline_up: antispoof_ingress -> nat44_egress -> qos_ingress_prog, the QoS
 *                stage keyed on the PRE-NAT source address (it runs on a copy
 *                of the header taken before SNAT; qos_ingress_prog never
 *                writes the frame, bpf/qos_ratelimit.c:178-222).
 *   pipeline_tc: the order the reference's TC hooks give (pkg/{antispoof,qos,nat}/tc_linux.go):
 *                antispoof_ingress -> qos_ingress_prog -> nat44_egress. */
static int run_pipeline(int up, ora_pkt *p) {
    int v = fn_antispoof(p);
    if (v == 2) return 2;
    if (up) {
        uint32_t hl = p->len < 64 ? p->len : 64;
        memcpy(g_scratch, p->data, hl);
        v = fn_nat_egress(p);
        if (v == 2) return 2;
        ora_pkt q = *p;
        q.data = g_scratch;
        v = fn_qos_ingress(&q);
        return v == 2 ? 2 : 0;
    }
    v = fn_qos_ingress(p);
    if (v == 2) return 2;
    v = fn_nat_egress(p);
    return v == 2 ? 2 : 0;
}

int ora_prog_run(int prog, ora_batch *b) {
    ensure_init();
    if (prog < 0 || prog > pipe_tc_id) return -EINVAL;
    g_now = b->now_ns;
    ora_prog_fn fn = prog < g_nprogs ? g_progs[prog].fn : NULL;
    for (uint32_t i = 0; i < b->n; i++) {
        ora_pkt p;
        if (b->now_v) g_now = b->now_v[i];
        p.data = b->pkts + (b->off16 ? (size_t)b->off16[i] * 16 : (size_t)i * b->stride);
        p.len = b->len[i];
        p.priority = b->priority ? b->priority[i] : 0;
        p.ctx_cookie = NULL;
        int v = fn ? fn(&p) : run_pipeline(prog == pipe_up_id, &p);
        b->verdict[i] = (uint8_t)v;
        b->len[i] = p.len;
        if (b->priority) b->priority[i] = p.priority;
    }
    return 0;
}
