/* TEST INFRASTRUCTURE — userspace stand-in for libbpf's <bpf/bpf_endian.h>
 * (little-endian host, as on x86-64 and as the eBPF target the reference
 * builds for). */
#ifndef BNG_ORACLE_SHIM_BPF_ENDIAN_H
#define BNG_ORACLE_SHIM_BPF_ENDIAN_H

#define bpf_htons(x) ((__u16)__builtin_bswap16((__u16)(x)))
#define bpf_ntohs(x) ((__u16)__builtin_bswap16((__u16)(x)))
#define bpf_htonl(x) ((__u32)__builtin_bswap32((__u32)(x)))
#define bpf_ntohl(x) ((__u32)__builtin_bswap32((__u32)(x)))

#endif
