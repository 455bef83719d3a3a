// bng_b200 — host-side sharding of one subscriber population over N dataplane contexts (one per GPU).
//
// every mutable table is keyed by something one subscriber owns, so frames and state
// partition by  shard = splitmix64(mac_key) % N  (bng_shard_of_mac) with no data-path collective.  What the
// control plane needs on the host is
//   * the MAC <-> IP directory it already has at lease time (reference pkg/dhcp/server.go:1062-1075 hands both to
//     the fast-path cache) — per-IP maps (subscriber_nat, qos_ingress/egress, nat_sessions, eim_table) follow the
//     subscriber's MAC to its shard;
//   * for the downstream direction (§8f-1) the (public IP, port block) -> shard table.  Blocks are laid out
//     deterministically by AllocateNAT (pkg/nat/manager.go:433-434: port_start = range_start + k * ports_per_sub),
//     so the owner of an inbound (dst ip, dst port) is one array lookup;
//   * replication of the small read-mostly maps (ip_pools, server_config, nat_config_map, alg_ports, hairpin_ips,
//     antispoof_config, allowed_ranges_v4, nat_pool, nat_private_ranges) to every shard.
// Router::Update / Lookup / Delete are what a cgo shim binds the Go managers' Map.Put / Lookup / Delete to when
// more than one GPU is in use; Steer* is what the ingest path (RSS / flow steering) implements in hardware.
#pragma once
#include <errno.h>
#include <string.h>

#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/bng_b200.h"
#include "bng_host.hpp"

namespace bng {
namespace shard {

enum class Route { ByMAC, ByPrivateIP, BySessionKey, ByEIMKey, ByReverseKey, Replicated };

inline Route RouteOf(const std::string &map) {
    if (map == "subscriber_bindings" || map == "subscriber_pools") return Route::ByMAC;
    if (map == "subscriber_nat" || map == "qos_ingress" || map == "qos_egress") return Route::ByPrivateIP;
    if (map == "nat_sessions") return Route::BySessionKey; // struct nat_key: src_ip = the subscriber's address
    if (map == "eim_table") return Route::ByEIMKey;        // struct eim_key: internal_ip
    if (map == "nat_reverse") return Route::ByReverseKey;  // struct nat_key: dst_ip/dst_port = public address, port
    return Route::Replicated;
}

// The MAC <-> IP directory and the public (address, port block) table.  Addresses are the 4 key bytes exactly as
// the maps hold them (memcpy'd into a u32), ports host order.
class Directory {
  public:
    explicit Directory(uint32_t world, uint16_t range_start = 1024, uint16_t ports_per_sub = 1024)
        : world_(world ? world : 1), range_start_(range_start), pps_(ports_per_sub ? ports_per_sub : 1024) {}
    uint32_t World() const { return world_; }
    uint32_t ShardOfMAC(uint64_t mac) const { return bng_shard_of_mac(mac, world_); }
    void Learn(uint64_t mac, uint32_t ip_key) { // lease granted: pkg/dhcp/server.go:1062-1075
        std::lock_guard<std::mutex> g(mu_);
        auto old = mac_ip_.find(mac);
        if (old != mac_ip_.end()) ip_mac_.erase(old->second);
        mac_ip_[mac] = ip_key;
        ip_mac_[ip_key] = mac;
    }
    void Forget(uint64_t mac) {
        std::lock_guard<std::mutex> g(mu_);
        auto it = mac_ip_.find(mac);
        if (it == mac_ip_.end()) return;
        ip_mac_.erase(it->second);
        mac_ip_.erase(it);
    }
    std::optional<uint32_t> ShardOfIP(uint32_t ip_key) const {
        std::lock_guard<std::mutex> g(mu_);
        auto it = ip_mac_.find(ip_key);
        if (it == ip_mac_.end()) return std::nullopt;
        return bng_shard_of_mac(it->second, world_);
    }
    // AllocateNAT gave `private_ip_key` the block [port_start, port_end] of public_ip_key
    void AddBlock(uint32_t public_ip_key, uint16_t port_start, uint32_t private_ip_key) {
        std::lock_guard<std::mutex> g(mu_);
        auto &v = blocks_[public_ip_key];
        size_t k = (size_t)(port_start - range_start_) / pps_;
        if (v.size() <= k) v.resize(k + 1, kNone);
        v[k] = private_ip_key;
    }
    void RemoveBlock(uint32_t public_ip_key, uint16_t port_start) {
        std::lock_guard<std::mutex> g(mu_);
        auto it = blocks_.find(public_ip_key);
        if (it == blocks_.end()) return;
        size_t k = (size_t)(port_start - range_start_) / pps_;
        if (k < it->second.size()) it->second[k] = kNone;
    }
    // owner of an inbound frame addressed to (public ip, port): the subscriber holding that block
    std::optional<uint32_t> ShardOfPublic(uint32_t public_ip_key, uint16_t port) const {
        uint32_t priv;
        {
            std::lock_guard<std::mutex> g(mu_);
            auto it = blocks_.find(public_ip_key);
            if (it == blocks_.end() || port < range_start_) return std::nullopt;
            size_t k = (size_t)(port - range_start_) / pps_;
            if (k >= it->second.size() || it->second[k] == kNone) return std::nullopt;
            priv = it->second[k];
        }
        return ShardOfIP(priv);
    }
    // ---- frame steering (what the NIC's flow steering does in front of the GPUs) ----
    static uint64_t MacKey(const uint8_t *m) {
        uint64_t k = 0;
        for (int i = 0; i < 6; i++) k = (k << 8) | m[i];
        return k;
    }
    uint32_t SteerUpstream(const uint8_t *frame, uint32_t len) const { // by source MAC
        return len >= 12 ? ShardOfMAC(MacKey(frame + 6)) : 0;
    }
    // by destination (public address, port / echo id); frames that are not translatable IPv4 go to `fallback`
    uint32_t SteerDownstream(const uint8_t *f, uint32_t len, uint32_t fallback = 0) const {
        if (len < 34 || f[12] != 0x08 || f[13] != 0x00) return fallback;
        uint32_t l4 = 14 + (uint32_t)(f[14] & 0x0f) * 4, daddr;
        memcpy(&daddr, f + 30, 4);
        uint32_t proto = f[23], poff;
        if (proto == 6 || proto == 17)
            poff = l4 + 2;
        else if (proto == 1)
            poff = l4 + 4;
        else
            return fallback;
        if (poff + 2 > len) return fallback;
        uint16_t port = (uint16_t)((f[poff] << 8) | f[poff + 1]);
        auto s = ShardOfPublic(daddr, port);
        return s ? *s : fallback;
    }

  private:
    static constexpr uint32_t kNone = 0xFFFFFFFFu;
    uint32_t world_;
    uint16_t range_start_, pps_;
    mutable std::mutex mu_;
    std::unordered_map<uint64_t, uint32_t> mac_ip_;
    std::unordered_map<uint32_t, uint64_t> ip_mac_;
    std::unordered_map<uint32_t, std::vector<uint32_t>> blocks_; // public ip -> private ip per block index
};

// N contexts behind one map API.
class Router {
  public:
    Router(std::vector<std::shared_ptr<Backend>> shards, std::shared_ptr<Directory> dir)
        : shards_(std::move(shards)), dir_(std::move(dir)) {}
    size_t World() const { return shards_.size(); }
    Backend &Shard(size_t i) { return *shards_[i]; }
    Directory &Dir() { return *dir_; }

    // -1: replicated, -ENOENT: the owner is not known (the address was never Learn()ed)
    int Owner(const std::string &map, const void *key) const {
        const uint8_t *k = (const uint8_t *)key;
        uint32_t ip;
        switch (RouteOf(map)) {
        case Route::ByMAC: {
            uint64_t mac;
            memcpy(&mac, k, 8);
            return (int)dir_->ShardOfMAC(mac);
        }
        case Route::ByPrivateIP:
        case Route::BySessionKey:
        case Route::ByEIMKey: {
            memcpy(&ip, k, 4);
            auto s = dir_->ShardOfIP(ip);
            return s ? (int)*s : -ENOENT;
        }
        case Route::ByReverseKey: {
            memcpy(&ip, k + 4, 4); // nat_key.dst_ip = the public address
            uint16_t port = (uint16_t)((k[10] << 8) | k[11]);
            auto s = dir_->ShardOfPublic(ip, port);
            return s ? (int)*s : -ENOENT;
        }
        default: return -1;
        }
    }
    int Update(const char *map, const void *key, const void *value, uint64_t flags = BNG_ANY, bool staged = false) {
        int o = Owner(map, key);
        if (o == -ENOENT) return o;
        int rc = 0;
        for (size_t i = 0; i < shards_.size(); i++) {
            if (o >= 0 && (size_t)o != i) continue;
            bng_ctx *c = shards_[i]->ctx;
            int id = bng_map_id(c, map);
            if (id < 0) return id;
            int r = staged && flags == BNG_ANY ? bng_map_update_staged(c, id, key, value) : bng_map_update(c, id, key, value, flags);
            if (r && !rc) rc = r;
        }
        return rc;
    }
    int Lookup(const char *map, const void *key, void *value_out) {
        int o = Owner(map, key);
        if (o == -ENOENT) return o;
        bng_ctx *c = shards_[o < 0 ? 0 : (size_t)o]->ctx; // replicated maps: any copy
        int id = bng_map_id(c, map);
        return id < 0 ? id : bng_map_lookup(c, id, key, value_out);
    }
    int Delete(const char *map, const void *key) {
        int o = Owner(map, key);
        if (o == -ENOENT) return o;
        int rc = 0;
        for (size_t i = 0; i < shards_.size(); i++) {
            if (o >= 0 && (size_t)o != i) continue;
            bng_ctx *c = shards_[i]->ctx;
            int id = bng_map_id(c, map);
            if (id < 0) return id;
            int r = bng_map_delete(c, id, key);
            if (r && !rc) rc = r;
        }
        return rc;
    }
    // PERCPU-style totals: the packed counter vector summed over the shards (host-side sum; with a communicator
    // per context bng_sync_reduce() does the same on the devices)
    int Totals(uint64_t out[BNG_NUM_STATS]) {
        memset(out, 0, sizeof(uint64_t) * BNG_NUM_STATS);
        for (auto &s : shards_) {
            uint64_t v[BNG_NUM_STATS];
            int r = bng_sync_reduce(s->ctx, v);
            if (r) return r;
            for (int i = 0; i < BNG_NUM_STATS; i++) out[i] += v[i];
        }
        return 0;
    }

  private:
    std::vector<std::shared_ptr<Backend>> shards_;
    std::shared_ptr<Directory> dir_;
};

} // namespace shard
} // namespace bng
