// BASELINE.json config #1: 1 000 synthetic DHCP DISCOVERs through the slow path, 256 of the
// clients already holding a lease, the rest allocating from a /22 — CPU only, one thread (the reference
// serialises allocation on the pool mutex, pkg/dhcp/pool.go:147).  Times the C++ restatement in
// bng_dhcp_slow.hpp; the Go original cannot be built here (no Go toolchain), which bench.py says next to
// the number.  Prints one JSON object.
//   g++ -std=c++17 -O2 dhcp_slow_bench.cpp -o dhcp_slow_bench   (no GPU, no libbng_b200.so needed at run time
//   beyond the header-only ABI declarations: the server runs with a null loader)
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "bng_dhcp_slow.hpp"

using namespace bng;

static uint64_t splitmix64(uint64_t &s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(int argc, char **argv) {
    const int n_req = 1000, n_leased = 256;
    const int rounds = argc > 1 ? atoi(argv[1]) : 2000;
    // inputs: DISCOVER of subscriber i = 0..999 (test/load/dhcp_benchmark.go's shape), built once
    std::vector<std::vector<uint8_t>> reqs;
    for (int i = 0; i < n_req; i++) reqs.push_back(dhcp::ClientMessage(dhcp::Discover, (uint32_t)i, 0x10000000u + (uint32_t)i));
    // which 256 subscribers already hold a lease: drawn with splitmix64(0xB2000001)
    std::vector<int> leased;
    {
        uint64_t s = 0xB2000001ull;
        std::vector<bool> pick(n_req, false);
        while ((int)leased.size() < n_leased) {
            int k = (int)(splitmix64(s) % n_req);
            if (!pick[k]) pick[k] = true, leased.push_back(k);
        }
    }
    uint64_t checksum = 0xcbf29ce484222325ull, offers = 0, bytes = 0;
    double seconds = 0;
    for (int r = 0; r < rounds; r++) {
        // fresh state per round (untimed): the pool, and the 256-entry lease map
        dhcp::PoolConfig c;
        c.ID = 1, c.Name = "residential", c.Network = "10.0.0.0/22", c.Gateway = "10.0.0.1", c.DNSServers = {"8.8.8.8", "8.8.4.4"};
        c.LeaseTimeSec = 3600;
        dhcp::PoolManager pm;
        pm.AddPool(*dhcp::Pool::New(c));
        int64_t now = 1700000000;
        dhcp::Server srv(0x0A000001u, &pm, nullptr, [&] { return now; });
        for (size_t j = 0; j < leased.size(); j++) {
            dhcp::Lease l;
            l.MAC = 0x020000000000ull | (uint32_t)leased[j];
            l.IP = 0x0A000300u + (uint32_t)j; // 10.0.3.x: inside the /22, apart from what the pool hands out first
            l.PoolID = 1, l.ExpiresAt = now + 1800;
            srv.InstallLease(l);
        }
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < n_req; i++) {
            auto out = srv.HandleDHCP(reqs[i].data(), reqs[i].size());
            if (!out.ok()) {
                fprintf(stderr, "request %d: %s\n", i, out.err.what().c_str());
                return 1;
            }
            bytes += out->size();
            if (r == 0)
                for (uint8_t b : *out) checksum = (checksum ^ b) * 0x100000001b3ull;
        }
        seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        offers += srv.offersTotal;
    }
    printf("{\"requests\": %llu, \"seconds\": %.6f, \"req_per_s\": %.1f, \"offers\": %llu, \"reply_bytes\": %llu, "
           "\"replies_fnv1a\": \"%016llx\", \"threads\": 1}\n",
           (unsigned long long)rounds * n_req, seconds, rounds * (double)n_req / seconds, (unsigned long long)offers,
           (unsigned long long)bytes, (unsigned long long)checksum);
    return 0;
}
