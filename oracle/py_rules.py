"""Second, independently written statement of Google snappy's compressor rules (pure Python, small inputs).

TEST INFRASTRUCTURE ONLY (see oracle/snappy_oracle.h).  Checks sjo_compress_rules in snappy_oracle.c: the two were
written separately (this one first, as the experiment that found which rule set reproduces pyarrow's bundled Google
snappy byte for byte) and must agree on every input.  rules: 0 = Snappy.jl, 1 = libsnappy
<= 1.1.7, 2 = Google snappy >= 1.1.9; the differences between them are the five lines marked `rules`."""


def _varint(n):
    o = bytearray()
    while n >= 0x80:
        o.append((n & 0x7F) | 0x80)
        n >>= 7
    o.append(n)
    return bytes(o)


def _ld32(F, q):
    return int.from_bytes(F[q:q + 4].ljust(4, b"\0"), "little")


def _literal(out, b, rules):
    n = len(b) - 1
    if (len(b) <= 60) if rules else (len(b) < 60):           # rules: the 60-byte literal
        out.append(n << 2)
    else:
        tmp = bytearray()
        while n > 0:
            tmp.append(n & 255)
            n >>= 8
        out.append((59 + len(tmp)) << 2)
        out += tmp
    out += b


def _op(out, off, ln):
    if ln < 12 and off < 2048:
        out.append(1 + ((ln - 4) << 2) + ((off >> 3) & 0xE0))
        out.append(off & 255)
    else:
        u = 2 + ((ln - 1) << 2) + (off << 8)
        out += bytes([u & 255, (u >> 8) & 255, (u >> 16) & 255])


def _copy(out, off, ln):
    if ln >= 12:
        while ln >= 68:
            _op(out, off, 64)
            ln -= 64
        if ln > 64:
            _op(out, off, 60)
            ln -= 60
    _op(out, off, ln)


def _fragment(F, entries, rules, out):
    n = len(F)
    shift = 32 - (entries.bit_length() - 1)
    if rules == 2:                                           # rules: the bucket of a hash
        def H(q):
            return (((_ld32(F, q) * 0x1E35A7BD) & 0xFFFFFFFF) >> 17) & (entries - 1)
    else:
        def H(q):
            return ((_ld32(F, q) * 0x1E35A7BD) & 0xFFFFFFFF) >> shift
    T = [0] * entries
    lim = n - 15 if rules else n - 16                        # rules: ip_limit
    ip = ne = 0
    if n >= 15:
        done = False
        while not done:
            skip = 32
            ip += 1
            nh = H(ip)
            nip = ip
            while True:
                ip, h = nip, nh
                b = skip >> 5
                skip += b
                nip = ip + b
                if nip > lim:
                    done = True
                    break
                nh = H(nip)
                cand = T[h]
                T[h] = ip
                if _ld32(F, cand) == _ld32(F, ip):
                    break
            if done:
                break
            _literal(out, F[ne:ip], rules)
            while True:
                m = 4
                while ip + m < n and F[cand + m] == F[ip + m]:
                    m += 1
                _copy(out, ip - cand, m)
                ip += m
                ne = ip
                if ip >= lim:
                    done = True
                    break
                T[H(ip - 1)] = ip - 1
                c = H(ip)
                cand = T[c]
                T[c] = ip
                if _ld32(F, ip) != _ld32(F, cand):
                    break
    if ne < n:
        _literal(out, F[ne:n], rules)


def _entries(x, cap):
    e = 256
    while e < cap and e < x:
        e *= 2
    return e


def compress(data, rules):
    data = bytes(data)
    N = len(data)
    out = bytearray(_varint(N))
    cap = 32768 if rules == 2 else 16384                     # rules: the largest table
    for s in range(0, N, 65536):
        F = data[s:s + 65536]
        _fragment(F, _entries(len(F) if rules else N, cap), rules, out)   # rules: table per fragment
    return bytes(out)
