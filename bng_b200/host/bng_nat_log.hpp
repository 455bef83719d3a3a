// bng_nat_log.hpp — the consumer side of the NAT event path: records the kernels stage in
// the nat_log_rb ring are drained through the C ABI and turned into the reference's compliance log lines.
//
// Restates pkg/nat/logging.go:
//   NATLogEntry :28-48, LogFormat :17-25, Logger.LogFromBPF :293-312, bpfEventTypeToString :314-333,
//   protocolToString :335-346, addEntry / Flush :349-394 (buffer, flush when full), formatEntry :416-440,
//   formatSyslog :470-481, formatCSV :484-500, formatNEL :503-522.
// plus the loop the reference leaves as a TODO (Manager.readLogEvents, pkg/nat/manager.go:682-696: "read from
// the ring buffer ... placeholder"): nat::PumpLog() = bng_events_drain -> Logger.LogFromBPF.
// Not restated: file rotation / gzip / retention (:525-683) — the sink here is any std::ostream.
//
// Byte-level conventions kept from the Go code: addresses print through keyToIP (manager.go:830-834:
// BigEndian.PutUint32 of the u32 AS THE GO STRUCT HOLDS IT — i.e. the little-endian load of the record's
// wire-order bytes, so a record for 100.64.0.9 prints "9.0.64.100"; SURVEY.md §7.3-3 lists this class of
// quirk) and ports print as the u16 the record holds (network-order bytes loaded little-endian).  Set
// Logger::wire_order to print what was on the wire instead.
// JSON follows encoding/json: struct fields in declaration order with omitempty, time as RFC 3339 with
// nanoseconds trimmed; formatNEL marshals maps, whose keys encoding/json sorts.
#pragma once

#include <ctime>
#include <functional>
#include <ostream>
#include <sstream>

#include "bng_host.hpp"

namespace bng {
namespace nat {

enum class LogFormat { JSON, Syslog, CSV, NEL }; // "json" | "syslog" | "csv" | "nel"

struct NATLogEntry { // logging.go:28-48
    int64_t TimestampSec = 0;
    uint32_t TimestampNsec = 0;
    std::string EventType;
    uint32_t SubscriberID = 0;
    std::string PrivateIP;
    uint16_t PrivatePort = 0;
    std::string PublicIP;
    uint16_t PublicPort = 0;
    std::string Protocol;
    std::string DestIP;
    uint16_t DestPort = 0;
    std::string SessionID;
    int64_t Duration = 0;
    uint64_t BytesSent = 0, BytesRecv = 0, PacketsSent = 0, PacketsRecv = 0;
    bool IsHairpin = false;
};

struct LoggerConfig { // :95-112 (the fields that shape the output)
    bool Enabled = true;
    LogFormat Format = LogFormat::JSON;
    size_t BufferSize = 1000;
};

class Logger {
  public:
    using Now = std::function<void(int64_t *sec, uint32_t *nsec)>; // time.Now().UTC()
    Logger(const LoggerConfig &cfg, std::ostream *sink, Now now = nullptr) : cfg_(cfg), sink_(sink), now_(std::move(now)) {
        if (cfg_.BufferSize == 0) cfg_.BufferSize = 1000; // :132-135
    }
    bool wire_order = false;

    void LogFromBPF(const LogEntry &e) { // :293-312
        if (!cfg_.Enabled) return;
        NATLogEntry n;
        stamp(n);
        n.EventType = bpfEventTypeToString(e.EventType);
        n.SubscriberID = e.SubscriberID;
        n.PrivateIP = ip(e.PrivateIP), n.PrivatePort = port(e.PrivatePort);
        n.PublicIP = ip(e.PublicIP), n.PublicPort = port(e.PublicPort);
        n.DestIP = ip(e.DestIP), n.DestPort = port(e.DestPort);
        n.Protocol = protocolToString(e.Protocol);
        addEntry(n);
    }
    void Flush() { // :376-394
        std::vector<NATLogEntry> batch;
        {
            std::lock_guard<std::mutex> g(mu_);
            batch.swap(buffer_);
        }
        for (const auto &e : batch) {
            std::string line = formatEntry(e);
            if (sink_) sink_->write(line.data(), (std::streamsize)line.size());
            linesWritten++;
        }
    }
    size_t Buffered() {
        std::lock_guard<std::mutex> g(mu_);
        return buffer_.size();
    }
    uint64_t linesWritten = 0;

    static const char *bpfEventTypeToString(uint32_t t) { // :314-333
        switch (t) {
        case NATLogSessionCreate: return "session_create";
        case NATLogSessionDelete: return "session_delete";
        case NATLogPortBlockAssign: return "port_block_assign";
        case NATLogPortBlockRelease: return "port_block_release";
        case NATLogPortExhaustion: return "port_exhaustion";
        case NATLogHairpin: return "hairpin";
        case NATLogALGTrigger: return "alg_trigger";
        default: return "unknown";
        }
    }
    static std::string protocolToString(uint8_t p) { // :335-346
        switch (p) {
        case 6: return "tcp";
        case 17: return "udp";
        case 1: return "icmp";
        default: return "proto_" + std::to_string((int)p);
        }
    }

    std::string formatEntry(const NATLogEntry &e) const { // :416-440
        switch (cfg_.Format) {
        case LogFormat::Syslog: return formatSyslog(e);
        case LogFormat::CSV: return formatCSV(e);
        case LogFormat::NEL: return formatNEL(e);
        default: return formatJSON(e);
        }
    }

  private:
    void stamp(NATLogEntry &n) {
        if (now_) {
            now_(&n.TimestampSec, &n.TimestampNsec);
        } else {
            timespec ts;
            clock_gettime(CLOCK_REALTIME, &ts);
            n.TimestampSec = ts.tv_sec, n.TimestampNsec = (uint32_t)ts.tv_nsec;
        }
    }
    void addEntry(const NATLogEntry &e) { // :349-361: flush when the buffer fills
        bool full;
        {
            std::lock_guard<std::mutex> g(mu_);
            buffer_.push_back(e);
            full = buffer_.size() >= cfg_.BufferSize;
        }
        if (full) Flush();
    }
    std::string ip(uint32_t field) const { // keyToIP(field).String()
        uint32_t v = wire_order ? __builtin_bswap32(field) : field;
        char b[20];
        snprintf(b, sizeof b, "%u.%u.%u.%u", v >> 24, (v >> 16) & 255, (v >> 8) & 255, v & 255);
        return b;
    }
    uint16_t port(uint16_t field) const { return wire_order ? (uint16_t)((field << 8) | (field >> 8)) : field; }
    static std::string rfc3339(int64_t sec, uint32_t nsec, bool nano) { // time.RFC3339 / RFC3339Nano, UTC
        time_t t = (time_t)sec;
        tm g;
        gmtime_r(&t, &g);
        char b[40];
        strftime(b, sizeof b, "%Y-%m-%dT%H:%M:%S", &g);
        std::string s = b;
        if (nano && nsec) {
            char f[16];
            snprintf(f, sizeof f, ".%09u", nsec);
            std::string frac = f;
            while (frac.back() == '0') frac.pop_back();
            s += frac;
        }
        return s + "Z";
    }
    std::string formatJSON(const NATLogEntry &e) const { // json.Marshal(entry) + '\n'
        std::ostringstream o;
        o << "{\"timestamp\":\"" << rfc3339(e.TimestampSec, e.TimestampNsec, true) << "\",\"event_type\":\"" << e.EventType << "\"";
        if (e.SubscriberID) o << ",\"subscriber_id\":" << e.SubscriberID;
        o << ",\"private_ip\":\"" << e.PrivateIP << "\"";
        if (e.PrivatePort) o << ",\"private_port\":" << e.PrivatePort;
        o << ",\"public_ip\":\"" << e.PublicIP << "\"";
        if (e.PublicPort) o << ",\"public_port\":" << e.PublicPort;
        if (!e.Protocol.empty()) o << ",\"protocol\":\"" << e.Protocol << "\"";
        if (!e.DestIP.empty()) o << ",\"dest_ip\":\"" << e.DestIP << "\"";
        if (e.DestPort) o << ",\"dest_port\":" << e.DestPort;
        if (!e.SessionID.empty()) o << ",\"session_id\":\"" << e.SessionID << "\"";
        if (e.Duration) o << ",\"duration_ms\":" << e.Duration;
        if (e.BytesSent) o << ",\"bytes_sent\":" << e.BytesSent;
        if (e.BytesRecv) o << ",\"bytes_recv\":" << e.BytesRecv;
        if (e.PacketsSent) o << ",\"packets_sent\":" << e.PacketsSent;
        if (e.PacketsRecv) o << ",\"packets_recv\":" << e.PacketsRecv;
        if (e.IsHairpin) o << ",\"is_hairpin\":true";
        o << "}\n";
        return o.str();
    }
    std::string formatSyslog(const NATLogEntry &e) const { // :470-481
        std::ostringstream o;
        o << rfc3339(e.TimestampSec, 0, false) << " NAT " << e.EventType << ": subscriber=" << e.SubscriberID << " private=" << e.PrivateIP
          << ":" << e.PrivatePort << " public=" << e.PublicIP << ":" << e.PublicPort << " dest=" << e.DestIP << ":" << e.DestPort
          << " proto=" << e.Protocol << " duration=" << e.Duration << "ms\n";
        return o.str();
    }
    std::string formatCSV(const NATLogEntry &e) const { // :484-500
        std::ostringstream o;
        o << rfc3339(e.TimestampSec, 0, false) << "," << e.EventType << "," << e.SubscriberID << "," << e.PrivateIP << "," << e.PrivatePort
          << "," << e.PublicIP << "," << e.PublicPort << "," << e.DestIP << "," << e.DestPort << "," << e.Protocol << "," << e.Duration
          << "," << e.BytesSent << "," << e.BytesRecv << "\n";
        return o.str();
    }
    std::string formatNEL(const NATLogEntry &e) const { // :503-522: maps marshal with sorted keys
        std::ostringstream o;
        o << "{\"age\":0,\"body\":{\"dest_ip\":\"" << e.DestIP << "\",\"dest_port\":" << e.DestPort << ",\"event\":\"" << e.EventType
          << "\",\"private_ip\":\"" << e.PrivateIP << "\",\"private_port\":" << e.PrivatePort << ",\"protocol\":\"" << e.Protocol
          << "\",\"public_ip\":\"" << e.PublicIP << "\",\"public_port\":" << e.PublicPort << ",\"subscriber\":" << e.SubscriberID
          << "},\"type\":\"NAT\"}\n";
        return o.str();
    }

    LoggerConfig cfg_;
    std::ostream *sink_;
    Now now_;
    std::vector<NATLogEntry> buffer_;
    std::mutex mu_;
};

// What Manager.readLogEvents would do once it reads the ring (manager.go:682-696): hand every staged record to
// the logger, in the order the programs emitted them (bng_events_drain sorts by batch and frame index).
inline size_t PumpLog(Manager &m, Logger &l, size_t max = 1 << 16) {
    size_t total = 0;
    for (;;) {
        auto recs = m.DrainLog(max);
        for (const auto &r : recs) l.LogFromBPF(r);
        total += recs.size();
        if (recs.size() < max) break;
    }
    return total;
}

} // namespace nat
} // namespace bng
