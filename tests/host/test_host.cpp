// Tests of the C++ host mirror (bng_b200/host/bng_host.hpp), written after the
// reference's own Go tests of the same types:
//   pkg/ebpf/loader_test.go (conversions :16-130, FNV-1a :221-249, nil-map errors :383-446, Close :989-1003)
//   test/ebpf/maps_test.go  (struct sizes :68-130, circuit-id key :235-269)
//   pkg/nat/manager_test.go (port blocks :164-247), pkg/nat/manager_additional_test.go (flags/log2 :395-477)
//   pkg/qos/manager_test.go (bookkeeping :80-222), pkg/antispoof/manager_test.go
// `test_host cpu` runs everything that needs no device (the reference's "never call Start/Load" mode);
// `test_host gpu` additionally loads the dataplane and round-trips through the C ABI.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../bng_b200/host/bng_dhcp_slow.hpp"
#include "../../bng_b200/host/bng_host.hpp"
#include "../../bng_b200/host/bng_nat_log.hpp"
#include "../../bng_b200/host/bng_shard.hpp"

using namespace bng;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        g_checks++;                                                              \
        if (!(cond)) {                                                           \
            g_fail++;                                                            \
            fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);      \
        }                                                                        \
    } while (0)
#define CHECK_EQ(a, b)                                                                                        \
    do {                                                                                                      \
        g_checks++;                                                                                           \
        auto va = (a);                                                                                        \
        auto vb = (b);                                                                                        \
        if (!(va == vb)) {                                                                                    \
            g_fail++;                                                                                         \
            fprintf(stderr, "FAIL %s:%d: %s == %s (%llu vs %llu)\n", __FILE__, __LINE__, #a, #b,              \
                    (unsigned long long)va, (unsigned long long)vb);                                          \
        }                                                                                                     \
    } while (0)
#define CHECK_ERR(e, text) CHECK((e).what() == std::string(text))
#define CHECK_STR(a, b)                                                                                    \
    do {                                                                                                   \
        g_checks++;                                                                                        \
        std::string sa = (a), sb = (b);                                                                    \
        if (sa != sb) {                                                                                    \
            g_fail++;                                                                                      \
            fprintf(stderr, "FAIL %s:%d: %s\n  got:  %s\n  want: %s\n", __FILE__, __LINE__, #a, sa.c_str(), sb.c_str()); \
        }                                                                                                  \
    } while (0)

static std::vector<uint8_t> bytes(const char *s) { return std::vector<uint8_t>(s, s + strlen(s)); }

static void test_conversions() {
    CHECK_EQ(ebpf::IPToUint32(IPv4(10, 0, 1, 1)), 0x0A000101u);
    CHECK_EQ(ebpf::IPToUint32(IPv4(192, 168, 1, 100)), 0xC0A80164u);
    CHECK_EQ(ebpf::IPToUint32(IPv4(255, 255, 255, 255)), 0xFFFFFFFFu);
    CHECK_EQ(ebpf::IPToUint32(IP{}), 0u);
    IP mapped(16, 0);
    mapped[10] = mapped[11] = 0xff;
    mapped[12] = 10, mapped[13] = 0, mapped[14] = 1, mapped[15] = 1;
    CHECK_EQ(ebpf::IPToUint32(mapped), 0x0A000101u);
    IP v6(16, 0x20);
    CHECK_EQ(ebpf::IPToUint32(v6), 0u);
    CHECK(ebpf::Uint32ToIP(0x0A000101) == IPv4(10, 0, 1, 1));
    CHECK_EQ(ebpf::MACToUint64(MAC{0xaa, 0xbb, 0xcc, 0xdd, 0xee, 0xff}), 0xAABBCCDDEEFFull);
    CHECK_EQ(ebpf::MACToUint64(MAC{0, 0, 0, 0, 0, 1}), 1ull);
    CHECK_EQ(ebpf::MACToUint64(MAC{1, 2, 3}), 0ull);
    CHECK(ebpf::Uint64ToMAC(0xAABBCCDDEEFFull) == (MAC{0xaa, 0xbb, 0xcc, 0xdd, 0xee, 0xff}));
    CHECK_EQ(ebpf::HashCircuitID({}), 0xcbf29ce484222325ull);
    CHECK_EQ(ebpf::HashCircuitID(bytes("a")), 0xaf63dc4c8601ec8cull);
    CHECK_EQ(ebpf::HashCircuitID(bytes("foobar")), 0x85944171f73967e8ull);
    CHECK(ebpf::HashCircuitID(bytes("eth 0/1/1:100")) != ebpf::HashCircuitID(bytes("eth 0/1/1:101")));
    auto k = ebpf::MakeCircuitIDKey(bytes("short"));
    CHECK(memcmp(k.b, "short", 5) == 0 && k.b[5] == 0 && k.b[31] == 0);
    std::vector<uint8_t> lng(40, 'x');
    k = ebpf::MakeCircuitIDKey(lng);
    CHECK(k.b[31] == 'x');
    uint64_t e = ebpf::LeaseExpiryFromDuration(std::chrono::seconds(3600));
    CHECK(e > 1700000000ull);
    CHECK_EQ(sizeof(ebpf::ServerConfig), 16u);
    CHECK_EQ(sizeof(ebpf::DHCPStats), 80u);
    CHECK_EQ(sizeof(ebpf::VLANKey), 4u);
    CHECK_EQ(sizeof(ebpf::CircuitIDKey), 32u);
    CHECK_EQ(sizeof(ebpf::PoolAssignment), 25u);
    CHECK_EQ(sizeof(ebpf::IPPool), 28u);
}

static void test_loader_unloaded() {
    auto bad = ebpf::Loader::NewLoader("");
    CHECK(!bad.ok());
    CHECK_ERR(bad.err, "interface name is required");
    auto lr = ebpf::Loader::NewLoader("eth0");
    CHECK(lr.ok());
    auto l = *lr.value;
    ebpf::PoolAssignment a;
    CHECK_ERR(l->AddSubscriber(1, a), "subscriber_pools map not loaded");
    CHECK_ERR(l->RemoveSubscriber(1), "subscriber_pools map not loaded");
    CHECK_ERR(l->GetSubscriber(1).err, "subscriber_pools map not loaded");
    CHECK_ERR(l->AddVLANSubscriber(1, 2, a), "vlan_subscriber_pools map not loaded");
    CHECK_ERR(l->GetVLANSubscriber(1, 2).err, "vlan_subscriber_pools map not loaded");
    CHECK(!l->HasVLANSupport());
    ebpf::IPPool p;
    CHECK_ERR(l->AddPool(1, p), "ip_pools map not loaded");
    CHECK_ERR(l->GetPool(1).err, "ip_pools map not loaded");
    CHECK_ERR(l->GetStats().err, "stats_map not loaded");
    CHECK_ERR(l->ResetStats(), "stats_map not loaded");
    CHECK_ERR(l->SetServerConfig(MAC(6, 1), IPv4(10, 0, 0, 1), 2), "server_config map not loaded");
    CHECK_ERR(l->GetServerConfig().err, "server_config map not loaded");
    CHECK_ERR(l->AddCircuitIDMapping(bytes("x"), 1), "circuit_id_map not loaded");
    CHECK_ERR(l->GetCircuitIDMapping(bytes("x")).err, "circuit_id_map not loaded");
    CHECK_ERR(l->CheckCircuitIDCollision(bytes("x"), 1).err, "circuit_id_map not loaded");
    CHECK_ERR(l->AddCircuitIDSubscriber(bytes("x"), a), "circuit_id_subscribers map not loaded");
    CHECK_ERR(l->GetCircuitIDSubscriber(bytes("x")).err, "circuit_id_subscribers map not loaded");
    CHECK(!l->HasCircuitIDSubscriberSupport());
    CHECK(!l->Close());
    CHECK(!l->Close()); // idempotent
}

static void test_nat_allocator() {
    CHECK(!nat::Manager::NewManager({}).ok());
    nat::ManagerConfig cfg;
    cfg.Interface = "eth0";
    cfg.EnableEIM = cfg.EnableEIF = cfg.EnableHairpin = cfg.EnableFTPALG = true;
    auto m = *nat::Manager::NewManager(cfg).value;
    CHECK_EQ(m->buildFlags(), 0x0Fu);
    CHECK_ERR(m->AddPublicIP(IP(16, 1)), "IPv4 address required");
    CHECK(!m->AddPublicIP(IPv4(203, 0, 113, 1)));
    CHECK(!m->AddPublicIPRange(IPv4(203, 0, 113, 2), IPv4(203, 0, 113, 3)));
    CHECK_ERR(m->AddPublicIPRange(IPv4(203, 0, 113, 9), IPv4(203, 0, 113, 3)), "start IP must be less than or equal to end IP");
    auto pool = m->GetPoolStats();
    CHECK_EQ(pool.size(), 3u);
    CHECK_EQ(pool[0].TotalPorts, 64512);
    CHECK_EQ(pool[0].MaxSubscribers, 63); // (65535-1024+1)/1024
    auto a1 = m->AllocateNAT(IPv4(100, 64, 0, 1));
    CHECK(a1.ok());
    CHECK_EQ(a1->PortStart, 1024);
    CHECK_EQ(a1->PortEnd, 2047);
    CHECK_EQ(a1->SubscriberID, 1u);
    CHECK(a1->PublicIP == IPv4(203, 0, 113, 1));
    auto a2 = m->AllocateNAT(IPv4(100, 64, 0, 2));
    CHECK_EQ(a2->PortStart, 2048);
    CHECK_EQ(a2->PortEnd, 3071);
    CHECK_EQ(a2->SubscriberID, 2u);
    auto again = m->AllocateNAT(IPv4(100, 64, 0, 1));
    CHECK_EQ(again->PortStart, 1024);
    CHECK_EQ(m->GetAllocationCount(), 2);
    for (int i = 3; i <= 63; i++) CHECK(m->AllocateNAT(ebpf::Uint32ToIP(0x64400000u + i)).ok());
    auto a64 = m->AllocateNAT(IPv4(100, 64, 0, 64)); // first IP is full: next pool entry, first block
    CHECK(a64->PublicIP == IPv4(203, 0, 113, 2));
    CHECK_EQ(a64->PortStart, 1024);
    CHECK_EQ(a64->PoolIndex, 1);
    CHECK(!m->DeallocateNAT(IPv4(100, 64, 0, 2)));
    CHECK(!m->DeallocateNAT(IPv4(100, 64, 0, 2))); // not allocated: nil
    CHECK(!m->GetAllocation(IPv4(100, 64, 0, 2)).has_value());
    CHECK_EQ(m->GetPoolStats()[0].Subscribers, 62);
    auto re = m->AllocateNAT(IPv4(100, 64, 9, 9)); // Subscribers-- makes the next block overlap a live one (kept)
    CHECK_EQ(re->PortStart, (uint16_t)(1024 + 62 * 1024));
    CHECK_EQ(re->SubscriberID, 65u);
    nat::ManagerConfig small;
    small.Interface = "eth0";
    small.PortsPerSubscriber = 32256;
    auto s = *nat::Manager::NewManager(small).value;
    s->AddPublicIP(IPv4(198, 51, 100, 1));
    CHECK(s->AllocateNAT(IPv4(10, 0, 0, 1)).ok());
    CHECK(s->AllocateNAT(IPv4(10, 0, 0, 2)).ok());
    CHECK_ERR(s->AllocateNAT(IPv4(10, 0, 0, 3)).err, "NAT pool exhausted: no available public IPs");
    CHECK_EQ(nat::log2(1), 0);
    CHECK_EQ(nat::log2(1024), 10);
    CHECK_EQ(nat::log2(1500), 10);
    CHECK_EQ(nat::log2(65536), 16);
    CHECK_ERR(m->ConfigureALG(21, 6, nat::ALGTypeFTP, true), "ALG map not loaded");
    CHECK_ERR(m->GetStats().err, "stats map not loaded");
    CHECK_ERR(m->GetEIMMapping(IPv4(10, 0, 0, 1), 80, 6).err, "EIM table not loaded");
    CHECK_ERR(m->LookupSession(IPv4(10, 0, 0, 1), IPv4(8, 8, 8, 8), 1, 2, 6).err, "sessions map not loaded");
    CHECK_EQ((uint32_t)nat::NATLogPortExhaustion, 5u);
    CHECK_EQ((uint32_t)nat::NATFlagPortContiguity, 0x40u);
}

static void test_qos_bookkeeping() {
    CHECK(!qos::Manager::NewManager({}).ok());
    qos::ManagerConfig cfg;
    cfg.Interface = "eth0";
    auto nopol = *qos::Manager::NewManager(cfg).value;
    CHECK_ERR(nopol->SetSubscriberPolicy(IPv4(10, 0, 0, 1), "guest"), "policy manager not configured");
    auto m = *qos::Manager::NewManager(cfg, qos::DefaultPolicies()).value;
    CHECK_ERR(m->SetSubscriberPolicy(IPv4(10, 0, 0, 1), "nope"), "policy not found: nope");
    CHECK(!m->SetSubscriberPolicy(IPv4(10, 0, 0, 1), "residential-100mbps"));
    CHECK(!m->SetSubscriberPolicy(IPv4(10, 0, 0, 2), "unlimited"));
    CHECK_EQ(m->GetSubscriberCount(), 2);
    qos::SubscriberQoS q;
    CHECK_ERR(m->SetSubscriberQoS(q), "subscriber IP required");
    q.Addr = IP(16, 1);
    CHECK_ERR(m->SetSubscriberQoS(q), "IPv4 address required");
    CHECK(!m->RemoveSubscriberQoS(IPv4(10, 0, 0, 2)));
    CHECK_EQ(m->GetSubscriberCount(), 1);
    CHECK_ERR(m->GetStats().err, "stats map not loaded");
    // burst defaulting (pkg/qos/manager.go:181-209)
    q.Addr = IPv4(10, 0, 0, 3);
    q.DownloadBPS = 100000000;
    q.UploadBPS = 20000000;
    CHECK_EQ(qos::Manager::egressBucket(q).BurstBytes, 10u * 1024 * 1024); // 12.5 MB capped at 10 MiB
    CHECK_EQ(qos::Manager::ingressBucket(q).BurstBytes, 2500000u);
    q.DownloadBPS = 100000;
    CHECK_EQ(qos::Manager::egressBucket(q).BurstBytes, 65536u);
    q.BurstBytes = 2000000;
    CHECK_EQ(qos::Manager::egressBucket(q).BurstBytes, 2000000u);
    CHECK_EQ(qos::Manager::ingressBucket(q).BurstBytes, 2500000u); // upload ignores BurstBytes
    CHECK_EQ(qos::Manager::egressBucket(q).Tokens, 2000000ull);
    CHECK_EQ(qos::Manager::egressBucket(q).LastUpdate, 0ull);
}

static void test_antispoof_bookkeeping() {
    CHECK(!antispoof::Manager::NewManager({}).ok());
    antispoof::ManagerConfig cfg;
    cfg.Interface = "eth0";
    cfg.DefaultMode = antispoof::ModeStrict;
    auto m = *antispoof::Manager::NewManager(cfg).value;
    CHECK_ERR(m->AddBinding(MAC{1, 2, 3}, IPv4(10, 0, 0, 1)), "invalid MAC address");
    CHECK(!m->AddBinding(MAC{2, 0, 0, 0, 0, 1}, IPv4(10, 0, 0, 1)));
    CHECK(!m->AddBinding(MAC{2, 0, 0, 0, 0, 2}, IPv4(10, 0, 0, 2)));
    CHECK_EQ(m->GetBindingCount(), 2);
    CHECK(!m->RemoveBinding(MAC{2, 0, 0, 0, 0, 2}));
    CHECK_EQ(m->GetBindingCount(), 1);
    CHECK_ERR(m->AddAllowedRange(IPv4(10, 0, 0, 0), 8), "ranges map not loaded");
    CHECK_ERR(m->GetStats().err, "stats map not loaded");
    CHECK(!m->SetMode(antispoof::ModeLoose));
    CHECK_EQ(sizeof(antispoof::SubscriberBinding), 24u);
}

// ---- with a device: everything round-trips through the C ABI ----
static std::shared_ptr<dhcp::Pool> make_pool(const char *net, const char *gw, int rs = 0, int re = 0, uint32_t id = 1);

static void test_gpu_roundtrips() {
    bng_open_opts o;
    memset(&o, 0, sizeof(o));
    o.struct_size = sizeof(o);
    o.device = -1;
    o.max_subscribers = 4096;
    o.max_nat_sessions = 8192;
    o.max_eim_mappings = 8192;
    o.max_batch = 4096;
    auto be = Backend::Open(&o);
    CHECK(be->ctx != nullptr);
    if (!be->ctx) {
        fprintf(stderr, "bng_open: %s\n", be->open_error.c_str());
        return;
    }
    be->wire_order_keys = true; // store addresses as the programs read them off the wire
    auto l = *ebpf::Loader::NewLoader("eth0", be).value;
    CHECK(!l->Load());
    CHECK(l->HasVLANSupport() && l->HasCircuitIDSubscriberSupport());
    ebpf::PoolAssignment a;
    a.PoolID = 7;
    a.AllocatedIP = 0x0100000A;
    a.LeaseExpiry = 1ull << 40;
    a.ClientClass = 2;
    CHECK(!l->AddSubscriber(0x020000000001ull, a));
    auto g = l->GetSubscriber(0x020000000001ull);
    CHECK(g.ok() && g->PoolID == 7 && g->LeaseExpiry == (1ull << 40) && g->ClientClass == 2);
    CHECK(!l->GetSubscriber(0x020000000002ull).ok());
    CHECK(!l->RemoveSubscriber(0x020000000001ull));
    CHECK((bool)l->RemoveSubscriber(0x020000000001ull)); // second delete: error, as Map.Delete does
    CHECK(!l->AddVLANSubscriber(100, 7, a));
    CHECK(l->GetVLANSubscriber(100, 7).ok() && !l->GetVLANSubscriber(100, 8).ok());
    ebpf::IPPool p;
    p.PrefixLen = 24;
    p.LeaseTime = 3600;
    CHECK(!l->AddPool(7, p));
    CHECK(l->GetPool(7).ok() && l->GetPool(7)->LeaseTime == 3600);
    CHECK(!l->SetServerConfig(MAC{2, 0xaa, 0xbb, 0xcc, 0xdd, 1}, IPv4(10, 0, 0, 1), 3));
    auto sc = l->GetServerConfig();
    CHECK(sc.ok() && sc->ServerMAC[1] == 0xaa && sc->InterfaceIndex == 3 && sc->ServerIP == 0x0100000Au);
    CHECK(!l->AddCircuitIDMapping(bytes("eth 0/1/1:100"), 42));
    CHECK(*l->GetCircuitIDMapping(bytes("eth 0/1/1:100")).value == 42ull);
    CHECK(*l->CheckCircuitIDCollision(bytes("eth 0/1/1:100"), 42).value == false);
    CHECK(*l->CheckCircuitIDCollision(bytes("eth 0/1/1:100"), 43).value == true);
    CHECK(*l->CheckCircuitIDCollision(bytes("other"), 43).value == false);
    CHECK(!l->AddCircuitIDSubscriber(bytes("eth 0/1/1:100"), a));
    CHECK(l->GetCircuitIDSubscriber(bytes("eth 0/1/1:100")).ok());
    auto st = l->GetStats();
    CHECK(st.ok() && st->TotalRequests == 0);

    antispoof::ManagerConfig ac;
    ac.Interface = "eth0";
    ac.DefaultMode = antispoof::ModeStrict;
    ac.Backend_ = be;
    auto am = *antispoof::Manager::NewManager(ac).value;
    CHECK(!am->Start());
    antispoof::Config cfgv;
    uint32_t zero = 0;
    CHECK_EQ(bng_map_lookup(be->ctx, be->Map("antispoof_config"), &zero, &cfgv), 0);
    CHECK(cfgv.DefaultMode == 1 && cfgv.LogViolations == 1);
    MAC mac{2, 0, 0, 0, 0, 9};
    CHECK(!am->AddBinding(mac, IPv4(100, 64, 0, 9)));
    IP v6(16, 0);
    v6[0] = 0x20, v6[15] = 9;
    CHECK(!am->AddBindingV6(mac, v6));
    antispoof::SubscriberBinding sb;
    uint64_t mk = ebpf::MACToUint64(mac);
    CHECK_EQ(bng_map_lookup(be->ctx, be->Map("subscriber_bindings"), &mk, &sb), 0);
    CHECK(sb.IPv4Valid == 1 && sb.IPv6Valid == 1 && sb.IPv6Addr[15] == 9 && sb.Mode == 1);
    CHECK_EQ(sb.IPv4Addr, 0x09004064u); // wire order 100.64.0.9 read little-endian
    CHECK(!am->AddAllowedRange(IPv4(100, 64, 0, 0), 10));

    qos::ManagerConfig qc;
    qc.Interface = "eth0";
    qc.Backend_ = be;
    auto qm = *qos::Manager::NewManager(qc, qos::DefaultPolicies()).value;
    CHECK(!qm->Start());
    CHECK(!qm->SetSubscriberPolicy(IPv4(100, 64, 0, 9), "residential-100mbps"));
    qos::TokenBucket tb;
    uint32_t ipk = 0x09004064u;
    CHECK_EQ(bng_map_lookup(be->ctx, be->Map("qos_egress"), &ipk, &tb), 0);
    CHECK(tb.RateBPS == 100000000ull && tb.BurstBytes == 2000000u && tb.Tokens == 2000000ull && tb.Priority == 4);
    CHECK_EQ(bng_map_lookup(be->ctx, be->Map("qos_ingress"), &ipk, &tb), 0);
    CHECK(tb.RateBPS == 20000000ull && tb.BurstBytes == 2500000u);

    nat::ManagerConfig nc;
    nc.Interface = "eth0";
    nc.EnableEIM = nc.EnableEIF = nc.EnableHairpin = nc.EnableFTPALG = true;
    nc.Backend_ = be;
    auto nm = *nat::Manager::NewManager(nc).value;
    CHECK(!nm->Start());
    CHECK(!nm->AddPublicIP(IPv4(203, 0, 113, 1)));
    auto al = nm->AllocateNAT(IPv4(100, 64, 0, 9));
    CHECK(al.ok());
    nat::SubscriberNAT sn;
    CHECK_EQ(bng_map_lookup(be->ctx, be->Map("subscriber_nat"), &ipk, &sn), 0);
    CHECK(sn.Block.PortStart == 1024 && sn.Block.PortEnd == 2047 && sn.Block.NextPort == 1024 && sn.Block.SubscriberID == 1 &&
          sn.Block.BlockSizeLog2 == 10);
    nat::NATConfig ncfg;
    CHECK_EQ(bng_map_lookup(be->ctx, be->Map("nat_config_map"), &zero, &ncfg), 0);
    CHECK(ncfg.Flags == 0x0F && ncfg.PortRangeStart == 1024 && ncfg.PortRangeEnd == 65535 && ncfg.DefaultPortsPerSub == 1024);
    nat::ALGConfig alg;
    uint32_t ak = (21u << 16) | 6;
    CHECK_EQ(bng_map_lookup(be->ctx, be->Map("alg_ports"), &ak, &alg), 0);
    CHECK(alg.Port == 21 && alg.ALGType == nat::ALGTypeFTP);

    // one frame through the pipeline: the managers' entries must match a real frame (wire-order keys)
    alignas(16) uint8_t f[64] = {0};
    const uint8_t hdr[] = {0x02, 0xff, 0xff, 0xff, 0xff, 0xfe, 2, 0, 0, 0, 0, 9, 0x08, 0x00, 0x45, 0, 0, 50, 0, 1, 0x40, 0, 64, 6,
                           0x12, 0x34, 100, 64, 0, 9, 8, 8, 8, 8, 0x9c, 0x40, 0x01, 0xbb};
    memcpy(f, hdr, sizeof(hdr));
    f[47] = 0x10;
    f[50] = 0xab, f[51] = 0xcd;
    uint32_t len = 64;
    uint8_t verdict = 9;
    bng_batch b;
    memset(&b, 0, sizeof(b));
    b.pkts = f;
    b.len = &len;
    b.verdict = &verdict;
    b.n = 1;
    b.stride = 64;
    b.now_ns = 1000000000ull;
    b.mem = BNG_MEM_HOST;
    CHECK_EQ(bng_prog_run(be->ctx, bng_prog_id(be->ctx, "pipeline_up"), &b), 0);
    CHECK_EQ((int)verdict, 0);
    CHECK(f[26] == 203 && f[27] == 0 && f[28] == 113 && f[29] == 1); // source rewritten to the public address
    auto as = am->GetStats();
    CHECK(as.ok() && as->PacketsAllowed == 1 && as->PacketsDropped == 0);
    auto qs = qm->GetStats();
    CHECK(qs.ok() && qs->PacketsPassed == 1 && qs->BytesPassed == 64);
    auto ns = nm->GetStats();
    CHECK(ns.ok() && ns->PacketsSNAT == 1 && ns->SessionsCreated == 1 && ns->EIMMisses == 1);
    auto ses = nm->LookupSession(IPv4(100, 64, 0, 9), IPv4(8, 8, 8, 8), 0x409c, 0xbb01, 6);
    CHECK(ses.ok() && ses->PacketsOut == 1 && ses->BytesOut == 64 && ses->Protocol == 6 && ses->OrigPort == 0x409c);
    auto eim = nm->GetEIMMapping(IPv4(100, 64, 0, 9), 0x409c, 6);
    CHECK(eim.ok() && eim->ExternalPort == 1024 && eim->RefCount == 1);
    auto log = nm->DrainLog();
    CHECK(log.size() == 1 && log[0].EventType == nat::NATLogSessionCreate && log[0].SubscriberID == 1);
    { // GPU ring -> host -> compliance log line: a second new flow of the same subscriber
        memcpy(f, hdr, sizeof(hdr));
        f[35] = 0x41; // source port 40001
        f[47] = 0x10, f[50] = 0xab, f[51] = 0xcd;
        b.now_ns = 2000000000ull;
        CHECK_EQ(bng_prog_run(be->ctx, bng_prog_id(be->ctx, "nat44_egress"), &b), 0);
        std::ostringstream sink;
        nat::LoggerConfig lc;
        lc.Format = nat::LogFormat::Syslog;
        nat::Logger lg(lc, &sink, [](int64_t *s, uint32_t *nsec) { *s = 1700000000, *nsec = 0; });
        lg.wire_order = true;
        CHECK_EQ(nat::PumpLog(*nm, lg), 1u);
        lg.Flush();
        CHECK_STR(sink.str(), std::string("2023-11-14T22:13:20Z NAT session_create: subscriber=1 private=100.64.0.9:40001 "
                                         "public=203.0.113.1:1025 dest=8.8.8.8:443 proto=tcp duration=0ms\n"));
        CHECK_EQ(nat::PumpLog(*nm, lg), 0u);
    }
    CHECK(!nm->DeallocateNAT(IPv4(100, 64, 0, 9)));
    CHECK(bng_map_lookup(be->ctx, be->Map("subscriber_nat"), &ipk, &sn) != 0);

    // ---- slow path -> fast path: what pkg/dhcp decides is what the GPU answers with from then on ----
    {
        dhcp::PoolManager pm(l.get());
        auto pool = make_pool("10.0.0.0/22", "10.0.0.1", 0, 0, 3);
        CHECK(!pm.AddPool(pool) && !pm.LastSyncError());
        auto gp = l->GetPool(3); // mirrored by AddPool (pkg/dhcp/pool.go:266-282)
        CHECK(gp.ok() && gp->Network == 0x0A000000u && gp->PrefixLen == 22 && gp->Gateway == 0x0A000001u &&
              gp->DNSPrimary == 0x08080808u && gp->DNSSecondary == 0x08080404u && gp->LeaseTime == 3600);
        CHECK(!l->SetServerConfig(MAC{2, 0xaa, 0, 0, 0, 1}, IPv4(10, 0, 0, 1), 2));
        int64_t t = 1700000000;
        dhcp::Server srv(0x0A000001u, &pm, l.get(), [&] { return t; });
        auto d = dhcp::ClientMessage(dhcp::Discover, 42, 0xABCD0001);
        auto offer = dhcp::Message::Parse(srv.HandleDHCP(d.data(), d.size())->data(), 300);
        CHECK(offer.ok() && offer->yiaddr == 0x0A000002u);
        CHECK(!l->GetSubscriber(0x02000000002Aull).ok()); // an OFFER does not populate the fast path
        auto rq = dhcp::ClientMessage(dhcp::Request, 42, 0xABCD0002, offer->yiaddr);
        auto ack = dhcp::Message::Parse(srv.HandleDHCP(rq.data(), rq.size())->data(), 300);
        CHECK(ack.ok() && ack->Type() == (int)dhcp::Ack && !srv.LastFastPathError());
        auto sub = l->GetSubscriber(0x02000000002Aull); // updateFastPathCache (server.go:1057-1075)
        CHECK(sub.ok() && sub->PoolID == 3 && sub->AllocatedIP == 0x0A000002u && sub->VlanID == 100 && sub->ClientClass == 1 &&
              sub->LeaseExpiry == (uint64_t)(t + 3600));
        // the client's next DISCOVER, as a frame on the wire, is answered by dhcp_fastpath_prog on the GPU
        std::vector<uint8_t> fr(362, 0);
        memset(&fr[0], 0xff, 6);
        memcpy(&fr[6], "\x02\x00\x00\x00\x00\x2a", 6);
        fr[12] = 0x08, fr[13] = 0x00;
        fr[14] = 0x45, fr[16] = (uint8_t)((362 - 14) >> 8), fr[17] = (uint8_t)(362 - 14), fr[22] = 64, fr[23] = 17;
        memset(&fr[30], 0xff, 4);
        fr[34] = 0, fr[35] = 68, fr[36] = 0, fr[37] = 67, fr[38] = (uint8_t)((362 - 34) >> 8), fr[39] = (uint8_t)(362 - 34);
        auto next = dhcp::ClientMessage(dhcp::Discover, 42, 0xABCD0003);
        memcpy(&fr[42], next.data(), 300);
        uint32_t flen = 362;
        uint8_t fv = 0xff;
        bng_batch fb;
        memset(&fb, 0, sizeof(fb));
        fb.pkts = fr.data(), fb.len = &flen, fb.verdict = &fv, fb.n = 1, fb.stride = 368, fb.now_ns = 5000000000ull, fb.mem = BNG_MEM_HOST;
        fr.resize(368);
        fb.pkts = fr.data();
        CHECK_EQ(bng_prog_run(be->ctx, bng_prog_id(be->ctx, "dhcp_fastpath_prog"), &fb), 0);
        CHECK_EQ((int)fv, 3); // XDP_TX: answered without the slow path
        CHECK_EQ((int)fr[42], 2); // BOOTREPLY
        uint32_t yi_stored = sub->AllocatedIP; // the program copies the stored word as is
        CHECK(memcmp(&fr[42 + 16], &yi_stored, 4) == 0);
        auto st = l->GetStats();
        CHECK(st.ok() && st->FastpathHits == 1);
    }
    CHECK(!l->Close());
}

// ---- DHCP slow path (bng_dhcp_slow.hpp), after pkg/dhcp/pool_test.go and the handleDiscover flow ----
static std::shared_ptr<dhcp::Pool> make_pool(const char *net, const char *gw, int rs, int re, uint32_t id) {
    dhcp::PoolConfig c;
    c.ID = id, c.Name = "test", c.Network = net, c.Gateway = gw, c.DNSServers = {"8.8.8.8", "8.8.4.4"};
    c.LeaseTimeSec = 3600, c.ReservedStart = rs, c.ReservedEnd = re, c.VlanID = 100, c.ClientClass = 1;
    auto p = dhcp::Pool::New(c);
    CHECK(p.ok());
    return *p;
}

static void test_dhcp_pool() {
    // TestNewPool (pool_test.go:9-42): /24 = 254 usable, minus 10 + 5 reserved; the gateway .1 is inside the reserved start
    auto p = make_pool("10.0.1.0/24", "10.0.1.1", 10, 5);
    CHECK_EQ(p->Stats().Available, 254 - 10 - 5);
    CHECK_EQ(p->SubnetMask, 0xFFFFFF00u);
    { // invalid inputs (NewPool :58-77)
        dhcp::PoolConfig c;
        c.Network = "10.0.1.0", c.Gateway = "10.0.1.1";
        CHECK(!dhcp::Pool::New(c).ok());
        c.Network = "10.0.1.0/24", c.Gateway = "nope";
        CHECK_ERR(dhcp::Pool::New(c).err, "invalid gateway IP: nope");
        c.Gateway = "10.0.1.1", c.DNSServers = {"8.8.8"};
        CHECK_ERR(dhcp::Pool::New(c).err, "invalid DNS server IP: 8.8.8");
    }
    // TestPoolAllocate (:44-83): first free address (gateway skipped), same MAC -> same address
    auto q = make_pool("192.168.1.0/24", "192.168.1.1");
    auto a = q->Allocate(0xAABBCCDDEEFFull);
    CHECK(a.ok() && *a == 0xC0A80102u && q->Contains(*a));
    auto a2 = q->Allocate(0xAABBCCDDEEFFull);
    CHECK(a2.ok() && *a2 == *a);
    CHECK_EQ(*q->Allocate(0xAABBCCDDEE00ull), 0xC0A80103u);
    // TestPoolRelease (:85-118): a /28 has 14 hosts, 13 without the gateway; release appends at the END
    auto r = make_pool("192.168.1.0/28", "192.168.1.1");
    int initial = r->Stats().Available;
    CHECK_EQ(initial, 13);
    auto ip = r->Allocate(1);
    CHECK_EQ(r->Stats().Available, initial - 1);
    r->Release(*ip);
    CHECK_EQ(r->Stats().Available, initial);
    CHECK_EQ(*r->Allocate(2), 0xC0A80103u); // .2 went to the back of the list
    // exhaustion
    for (uint64_t m = 10; r->Allocate(m).ok(); m++) {}
    CHECK_ERR(r->Allocate(999).err, "pool test exhausted");
    CHECK(!r->Contains(0xC0A80110u) && r->Contains(0xC0A8010Fu));
    // PoolManager (:191-260): first pool is the default; duplicates refused
    dhcp::PoolManager pm;
    CHECK(pm.ClassifyClient(1) == nullptr);
    CHECK(!pm.AddPool(p));
    CHECK_ERR(pm.AddPool(p), "pool 1 already exists");
    auto p2 = make_pool("10.0.2.0/24", "10.0.2.1", 0, 0, 2);
    CHECK(!pm.AddPool(p2));
    CHECK(pm.ClassifyClient(1) == p);
    CHECK(!pm.SetDefaultPool(2));
    CHECK(pm.ClassifyClient(1) == p2);
    CHECK_ERR(pm.SetDefaultPool(9), "pool 9 not found");
    CHECK(!pm.RemovePool(2));
    CHECK(pm.ClassifyClient(1) == p);
}

static void test_dhcp_slow_path() {
    int64_t t = 1700000000;
    dhcp::PoolManager pm;
    auto pool = make_pool("10.0.0.0/22", "10.0.0.1");
    CHECK(!pm.AddPool(pool));
    const uint32_t server_ip = 0x0A000001u;
    dhcp::Server srv(server_ip, &pm, nullptr, [&] { return t; });
    // DISCOVER of subscriber 5 -> OFFER with the pool's first address and the pool's parameters
    auto req = dhcp::ClientMessage(dhcp::Discover, 5, 0x11223344);
    CHECK_EQ(req.size(), 300u);
    auto out = srv.HandleDHCP(req.data(), req.size());
    CHECK(out.ok() && out->size() >= 300);
    auto offer = dhcp::Message::Parse(out->data(), out->size());
    CHECK(offer.ok());
    CHECK_EQ((int)offer->op, 2);
    CHECK_EQ(offer->xid, 0x11223344u);
    CHECK_EQ(offer->flags, 0x8000);
    CHECK_EQ(offer->Type(), (int)dhcp::Offer);
    CHECK_EQ(offer->yiaddr, 0x0A000002u); // .1 is the gateway
    CHECK_EQ(offer->siaddr, server_ip);
    CHECK(memcmp(offer->chaddr, "\x02\x00\x00\x00\x00\x05", 6) == 0);
    CHECK((offer->options[dhcp::OptServerID] == std::vector<uint8_t>{10, 0, 0, 1}));
    CHECK((offer->options[dhcp::OptLeaseTime] == std::vector<uint8_t>{0, 0, 0x0e, 0x10}));
    CHECK((offer->options[dhcp::OptSubnetMask] == std::vector<uint8_t>{255, 255, 252, 0}));
    CHECK((offer->options[dhcp::OptRouter] == std::vector<uint8_t>{10, 0, 0, 1}));
    CHECK((offer->options[dhcp::OptDNS] == std::vector<uint8_t>{8, 8, 8, 8, 8, 8, 4, 4}));
    CHECK_EQ(srv.offersTotal, 1u);
    // options leave in ascending code order and the message ends 255, padded to 300 bytes
    CHECK((*out)[240] == dhcp::OptSubnetMask && (*out)[246] == dhcp::OptRouter);
    // the same client asks again: Pool.Allocate returns the same address, no lease exists yet
    auto again = srv.HandleDHCP(req.data(), req.size());
    CHECK(dhcp::Message::Parse(again->data(), again->size())->yiaddr == 0x0A000002u);
    CHECK_EQ(srv.ActiveLeases(), 0u);
    // REQUEST for the offered address -> ACK, lease recorded with the pool's lease time
    auto rq = dhcp::ClientMessage(dhcp::Request, 5, 0x11223345, 0x0A000002u);
    auto ack = dhcp::Message::Parse(srv.HandleDHCP(rq.data(), rq.size())->data(), 300);
    CHECK(ack.ok() && ack->Type() == (int)dhcp::Ack && ack->yiaddr == 0x0A000002u);
    CHECK_EQ(srv.ActiveLeases(), 1u);
    // renewal with a different address -> NAK (:586-590); an address outside the pool for a new client -> NAK
    auto bad = dhcp::ClientMessage(dhcp::Request, 5, 1, 0x0A000009u);
    CHECK_EQ(dhcp::Message::Parse(srv.HandleDHCP(bad.data(), bad.size())->data(), 300)->Type(), (int)dhcp::Nak);
    auto outside = dhcp::ClientMessage(dhcp::Request, 6, 1, 0xC0A80001u);
    CHECK_EQ(dhcp::Message::Parse(srv.HandleDHCP(outside.data(), outside.size())->data(), 300)->Type(), (int)dhcp::Nak);
    CHECK_EQ(srv.naksTotal, 2u);
    // a live lease is reused by DISCOVER (:420-424); once expired the pool is asked again (same MAC -> same address)
    dhcp::Lease seeded;
    seeded.MAC = 0x020000000007ull, seeded.IP = 0x0A000155u, seeded.PoolID = 1, seeded.ExpiresAt = t + 100;
    srv.InstallLease(seeded);
    auto d7 = dhcp::ClientMessage(dhcp::Discover, 7, 9);
    CHECK_EQ(dhcp::Message::Parse(srv.HandleDHCP(d7.data(), d7.size())->data(), 300)->yiaddr, 0x0A000155u);
    t += 200;
    CHECK_EQ(dhcp::Message::Parse(srv.HandleDHCP(d7.data(), d7.size())->data(), 300)->yiaddr, 0x0A000003u);
    // malformed input
    CHECK(!srv.HandleDHCP(req.data(), 100).ok());
    // no pool at all
    dhcp::PoolManager empty;
    dhcp::Server none(server_ip, &empty, nullptr);
    CHECK_ERR(none.HandleDHCP(req.data(), req.size()).err, "no pool available for client");
}

// ---- NAT compliance log lines (bng_nat_log.hpp), formats of pkg/nat/logging.go:416-522 ----
static void test_nat_log_formats() {
    nat::LogEntry e;
    memset(&e, 0, sizeof(e));
    e.EventType = nat::NATLogSessionCreate;
    e.SubscriberID = 7;
    e.PrivateIP = 0x09004064u;  // bytes 64 40 00 09 = 100.64.0.9 on the wire, as the u32 the Go struct holds
    e.PublicIP = 0x017100CBu;   // 203.0.113.1
    e.DestIP = 0x08080808u;
    e.PrivatePort = 0x409C;     // bytes 9c 40 = 40000
    e.PublicPort = 0x0004;      // bytes 04 00 = 1024
    e.DestPort = 0xBB01;        // 443
    e.Protocol = 6;
    auto clock = [](int64_t *s, uint32_t *ns) { *s = 1700000000, *ns = 123450000; };
    auto line = [&](nat::LogFormat f, bool wire) {
        std::ostringstream sink;
        nat::LoggerConfig c;
        c.Format = f;
        nat::Logger l(c, &sink, clock);
        l.wire_order = wire;
        l.LogFromBPF(e);
        CHECK_EQ(l.Buffered(), 1u);
        l.Flush();
        CHECK_EQ(l.Buffered(), 0u);
        return sink.str();
    };
    // the reference's conventions: keyToIP(BigEndian) of the little-endian-loaded field, ports as loaded
    CHECK_STR(line(nat::LogFormat::JSON, false),
             std::string("{\"timestamp\":\"2023-11-14T22:13:20.12345Z\",\"event_type\":\"session_create\",\"subscriber_id\":7,"
                         "\"private_ip\":\"9.0.64.100\",\"private_port\":16540,\"public_ip\":\"1.113.0.203\",\"public_port\":4,"
                         "\"protocol\":\"tcp\",\"dest_ip\":\"8.8.8.8\",\"dest_port\":47873}\n"));
    // ... and what was on the wire
    CHECK_STR(line(nat::LogFormat::Syslog, true),
             std::string("2023-11-14T22:13:20Z NAT session_create: subscriber=7 private=100.64.0.9:40000 public=203.0.113.1:1024 "
                         "dest=8.8.8.8:443 proto=tcp duration=0ms\n"));
    CHECK_STR(line(nat::LogFormat::CSV, true),
             std::string("2023-11-14T22:13:20Z,session_create,7,100.64.0.9,40000,203.0.113.1,1024,8.8.8.8,443,tcp,0,0,0\n"));
    CHECK_STR(line(nat::LogFormat::NEL, true),
             std::string("{\"age\":0,\"body\":{\"dest_ip\":\"8.8.8.8\",\"dest_port\":443,\"event\":\"session_create\",\"private_ip\":"
                         "\"100.64.0.9\",\"private_port\":40000,\"protocol\":\"tcp\",\"public_ip\":\"203.0.113.1\",\"public_port\":1024,"
                         "\"subscriber\":7},\"type\":\"NAT\"}\n"));
    CHECK_STR(std::string(nat::Logger::bpfEventTypeToString(5)), std::string("port_exhaustion"));
    CHECK_STR(std::string(nat::Logger::bpfEventTypeToString(99)), std::string("unknown"));
    CHECK_STR(nat::Logger::protocolToString(47), std::string("proto_47"));
    { // omitempty: zero ports / subscriber vanish from the JSON; a full buffer flushes by itself; disabled = silent
        nat::LogEntry z;
        memset(&z, 0, sizeof(z));
        z.EventType = nat::NATLogALGTrigger, z.Protocol = 17;
        std::ostringstream sink;
        nat::LoggerConfig c;
        c.BufferSize = 2;
        nat::Logger l(c, &sink, [](int64_t *s, uint32_t *ns) { *s = 0, *ns = 0; });
        l.LogFromBPF(z);
        CHECK(sink.str().empty());
        l.LogFromBPF(z);
        CHECK_EQ(l.linesWritten, 2u);
        CHECK_STR(sink.str().substr(0, sink.str().find('\n') + 1),
                 std::string("{\"timestamp\":\"1970-01-01T00:00:00Z\",\"event_type\":\"alg_trigger\",\"private_ip\":\"0.0.0.0\","
                             "\"public_ip\":\"0.0.0.0\",\"protocol\":\"udp\",\"dest_ip\":\"0.0.0.0\"}\n"));
        c.Enabled = false;
        nat::Logger off(c, &sink);
        off.LogFromBPF(z);
        CHECK_EQ(off.Buffered(), 0u);
    }
}

// ---- sharding: directory logic needs no device ----
static uint32_t ipkey(uint8_t a, uint8_t b, uint8_t c, uint8_t d) { // the 4 address bytes as the maps hold them
    uint8_t v[4] = {a, b, c, d};
    uint32_t k;
    memcpy(&k, v, 4);
    return k;
}

static void test_shard_directory() {
    shard::Directory dir(4, 1024, 1024);
    for (uint32_t i = 0; i < 1000; i++) dir.Learn(0x020000000000ull + i, ipkey(100, 64, (uint8_t)(i >> 8), (uint8_t)i));
    uint32_t per[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < 1000; i++) {
        uint32_t s = dir.ShardOfMAC(0x020000000000ull + i);
        CHECK(s < 4);
        per[s]++;
        CHECK_EQ(*dir.ShardOfIP(ipkey(100, 64, (uint8_t)(i >> 8), (uint8_t)i)), s); // the IP follows its MAC
    }
    for (int s = 0; s < 4; s++) CHECK(per[s] > 150);
    CHECK(!dir.ShardOfIP(ipkey(10, 9, 9, 9)).has_value());
    // port blocks by the AllocateNAT rule (pkg/nat/manager.go:433-434): subscriber k of a public IP owns
    // [1024 + 1024 k, 1024 + 1024 (k + 1) - 1]
    uint32_t pub = ipkey(203, 0, 113, 1);
    for (uint32_t k = 0; k < 63; k++) dir.AddBlock(pub, (uint16_t)(1024 + 1024 * k), ipkey(100, 64, 0, (uint8_t)k));
    for (uint32_t k = 0; k < 63; k++) {
        uint32_t want = dir.ShardOfMAC(0x020000000000ull + k);
        CHECK_EQ(*dir.ShardOfPublic(pub, (uint16_t)(1024 + 1024 * k)), want);
        CHECK_EQ(*dir.ShardOfPublic(pub, (uint16_t)(1024 + 1024 * k + 1023)), want);
    }
    CHECK(!dir.ShardOfPublic(pub, 80).has_value());                // below the port range
    CHECK(!dir.ShardOfPublic(ipkey(203, 0, 113, 2), 2000).has_value()); // unknown public address
    dir.RemoveBlock(pub, 1024);
    CHECK(!dir.ShardOfPublic(pub, 1500).has_value());
    dir.Forget(0x020000000001ull);
    CHECK(!dir.ShardOfIP(ipkey(100, 64, 0, 1)).has_value());
    // routing classes of the reference's map names
    CHECK(shard::RouteOf("subscriber_pools") == shard::Route::ByMAC);
    CHECK(shard::RouteOf("qos_egress") == shard::Route::ByPrivateIP);
    CHECK(shard::RouteOf("nat_reverse") == shard::Route::ByReverseKey);
    CHECK(shard::RouteOf("ip_pools") == shard::Route::Replicated && shard::RouteOf("hairpin_ips") == shard::Route::Replicated);
}

static void put16be(uint8_t *p, uint16_t v) {
    p[0] = (uint8_t)(v >> 8);
    p[1] = (uint8_t)v;
}
// 64-byte Ethernet / IPv4 / UDP frame (checksums are updated incrementally by the programs: any start value does)
static void udp_frame(uint8_t *f, uint64_t smac, uint64_t dmac, uint32_t sip, uint32_t dip, uint16_t sport, uint16_t dport) {
    memset(f, 0, 64);
    for (int i = 0; i < 6; i++) f[i] = (uint8_t)(dmac >> (8 * (5 - i))), f[6 + i] = (uint8_t)(smac >> (8 * (5 - i)));
    f[12] = 0x08, f[14] = 0x45, f[22] = 64, f[23] = 17;
    put16be(f + 16, 50);
    memcpy(f + 26, &sip, 4);
    memcpy(f + 30, &dip, 4);
    put16be(f + 24, 0x1234);
    put16be(f + 34, sport);
    put16be(f + 36, dport);
    put16be(f + 38, 30);
    put16be(f + 40, 0x4321);
}

// Two shards (two contexts; on one device here): a flow's upstream frames land on its subscriber's shard by
// source MAC, its replies on the same shard by (public address, port block), and only there are they translated.
static void test_gpu_two_shards() {
    const uint32_t world = 2, n_subs = 96;
    std::vector<std::shared_ptr<Backend>> bes;
    for (uint32_t r = 0; r < world; r++) {
        bng_open_opts o;
        memset(&o, 0, sizeof(o));
        o.struct_size = sizeof(o);
        o.device = -1;
        o.max_subscribers = 1024, o.max_nat_sessions = 4096, o.max_eim_mappings = 4096, o.max_batch = 4096;
        o.rank = r, o.world = world;
        auto be = Backend::Open(&o);
        CHECK(be->ctx != nullptr);
        if (!be->ctx) return;
        bes.push_back(be);
    }
    auto dir = std::make_shared<shard::Directory>(world, 1024, 1024);
    shard::Router rt(bes, dir);
    antispoof::Config ac{};
    ac.DefaultMode = 1;
    uint32_t zero = 0;
    CHECK_EQ(rt.Update("antispoof_config", &zero, &ac), 0); // replicated
    nat::NATConfig nc{};
    nc.Flags = 0x0F;
    nc.PortRangeStart = 1024, nc.PortRangeEnd = 65535, nc.DefaultPortsPerSub = 1024;
    CHECK_EQ(rt.Update("nat_config_map", &zero, &nc), 0);
    const uint32_t pub0 = ipkey(203, 0, 113, 0);
    uint32_t owner_count[2] = {0, 0};
    for (uint32_t i = 0; i < n_subs; i++) {
        uint64_t mac = 0x020000000000ull + i;
        uint32_t ip = ipkey(100, 64, 0, (uint8_t)i);
        // nothing per-IP can be placed before the directory knows whose address it is
        nat::SubscriberNAT sn{};
        CHECK_EQ(rt.Update("subscriber_nat", &ip, &sn), -ENOENT);
        dir->Learn(mac, ip);
        antispoof::SubscriberBinding sb{};
        sb.IPv4Addr = ip, sb.IPv4Valid = 1, sb.Mode = 1;
        CHECK_EQ(rt.Update("subscriber_bindings", &mac, &sb, BNG_ANY, true), 0); // staged: applied at the batch boundary
        uint32_t pub = pub0 + ((i / 63) << 24); // next public address after 63 blocks (last octet, wire order)
        uint16_t ps = (uint16_t)(1024 + 1024 * (i % 63));
        sn.Block.PublicIP = pub, sn.Block.PortStart = ps, sn.Block.PortEnd = (uint16_t)(ps + 1023), sn.Block.NextPort = ps;
        sn.Block.SubscriberID = i + 1, sn.Block.BlockSizeLog2 = 10;
        CHECK_EQ(rt.Update("subscriber_nat", &ip, &sn, BNG_ANY, true), 0);
        dir->AddBlock(pub, ps, ip);
        owner_count[dir->ShardOfMAC(mac)]++;
    }
    CHECK(owner_count[0] > 20 && owner_count[1] > 20);
    // each entry lives on exactly one shard
    for (uint32_t i = 0; i < n_subs; i++) {
        uint32_t ip = ipkey(100, 64, 0, (uint8_t)i), own = dir->ShardOfMAC(0x020000000000ull + i);
        nat::SubscriberNAT sn{};
        for (uint32_t r = 0; r < world; r++) {
            int rc = bng_map_lookup(bes[r]->ctx, bes[r]->Map("subscriber_nat"), &ip, &sn);
            CHECK_EQ(rc, r == own ? 0 : -ENOENT);
        }
        CHECK_EQ(rt.Lookup("subscriber_nat", &ip, &sn), 0);
    }
    // ---- upstream: two frames per subscriber, steered by source MAC ----
    const uint32_t n = 2 * n_subs;
    std::vector<uint8_t> up(n * 64);
    std::vector<std::vector<uint32_t>> mine(world);
    for (uint32_t j = 0; j < n; j++) {
        uint32_t i = j / 2;
        udp_frame(&up[j * 64], 0x020000000000ull + i, 0x02fffffffffeull, ipkey(100, 64, 0, (uint8_t)i), ipkey(8, 8, 8, 8),
                  (uint16_t)(4000 + (j & 1)), 53);
        mine[dir->SteerUpstream(&up[j * 64], 64)].push_back(j);
    }
    int prog_up = bng_prog_id(bes[0]->ctx, "pipeline_up"), prog_in = bng_prog_id(bes[0]->ctx, "nat44_ingress");
    std::vector<uint8_t> translated(n * 64);
    for (uint32_t r = 0; r < world; r++) {
        std::vector<uint8_t> a(mine[r].size() * 64), v(mine[r].size());
        std::vector<uint32_t> len(mine[r].size(), 64);
        for (size_t k = 0; k < mine[r].size(); k++) memcpy(&a[k * 64], &up[mine[r][k] * 64], 64);
        bng_batch b{};
        b.pkts = a.data(), b.len = len.data(), b.verdict = v.data(), b.n = (uint32_t)mine[r].size(), b.stride = 64;
        b.now_ns = 1000000000ull, b.mem = BNG_MEM_HOST;
        CHECK_EQ(bng_prog_run(bes[r]->ctx, prog_up, &b), 0);
        for (size_t k = 0; k < mine[r].size(); k++) {
            CHECK_EQ(v[k], 0);
            memcpy(&translated[mine[r][k] * 64], &a[k * 64], 64);
        }
    }
    uint64_t tot[BNG_NUM_STATS];
    CHECK_EQ(rt.Totals(tot), 0);
    CHECK_EQ(tot[0], (uint64_t)n);  // antispoof allowed: every frame found its binding on its shard
    CHECK_EQ(tot[10], (uint64_t)n); // packets_snat
    CHECK_EQ(tot[15], (uint64_t)n); // sessions_created: one per (subscriber, port)
    // ---- downstream: the replies go to the owner by (public address, port block) ----
    for (uint32_t j = 0; j < n; j++) {
        const uint8_t *t = &translated[j * 64];
        uint32_t nat_ip, i = j / 2, own = dir->ShardOfMAC(0x020000000000ull + i);
        memcpy(&nat_ip, t + 26, 4);
        CHECK(nat_ip != ipkey(100, 64, 0, (uint8_t)i)); // it was translated
        uint8_t reply[64];
        udp_frame(reply, 0x02fffffffffeull, 0x020000000000ull + i, ipkey(8, 8, 8, 8), nat_ip, 53, (uint16_t)((t[34] << 8) | t[35]));
        CHECK_EQ(dir->SteerDownstream(reply, 64, 99), own);
        for (uint32_t r = 0; r < world; r++) { // translated on the owner, untouched anywhere else
            uint8_t fr[64], vd = 9;
            uint32_t l = 64, daddr;
            memcpy(fr, reply, 64);
            bng_batch b{};
            b.pkts = fr, b.len = &l, b.verdict = &vd, b.n = 1, b.stride = 64, b.now_ns = 2000000000ull, b.mem = BNG_MEM_HOST;
            CHECK_EQ(bng_prog_run(bes[r]->ctx, prog_in, &b), 0);
            memcpy(&daddr, fr + 30, 4);
            if (r == own) {
                CHECK_EQ(daddr, ipkey(100, 64, 0, (uint8_t)i));
                CHECK_EQ((uint32_t)((fr[36] << 8) | fr[37]), 4000u + (j & 1));
            } else {
                CHECK(!memcmp(fr, reply, 64));
            }
        }
    }
    CHECK_EQ(rt.Totals(tot), 0);
    CHECK_EQ(tot[11], (uint64_t)n); // packets_dnat
}

int main(int argc, char **argv) {
    std::string mode = argc > 1 ? argv[1] : "cpu";
    test_conversions();
    test_dhcp_pool();
    test_dhcp_slow_path();
    test_nat_log_formats();
    test_loader_unloaded();
    test_nat_allocator();
    test_qos_bookkeeping();
    test_antispoof_bookkeeping();
    test_shard_directory();
    if (mode == "gpu") test_gpu_roundtrips();
    if (mode == "gpu") test_gpu_two_shards();
    printf("%s: %d checks, %d failed\n", mode.c_str(), g_checks, g_fail);
    return g_fail ? 1 : 0;
}
