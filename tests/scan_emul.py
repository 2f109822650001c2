"""Python emulations of the device algorithms of csrc/tg_scan.cu (window-table coverage scan, set-algebra hits).

Test infrastructure: they restate, position by position, what the CUDA kernels compute (same tables, same closed forms,
same tile-free formulation), so the algorithms can be checked against the oracle on the CPU before any GPU time is spent.
Bit sets over k-mers are Python ints.
"""
W = 7                                   # window length of the lookup tables (tg_scan.cu: SC_W)
CODE = {'A': 0, 'C': 1, 'T': 2, 'G': 3}


def is_bordered(s):
    return any(s[d:] == s[:len(s) - d] for d in range(1, len(s)))


def window_code(seq, p):
    """2 bits per base, first base in the low bits; N (anything outside ACGT) packs as 0 like tg_scan_pack."""
    w = 0
    for t in range(W):
        c = seq[p + t] if p + t < len(seq) else 'N'
        w |= CODE.get(c, 0) << (2 * t)
    return w


def window_str(w):
    return ''.join('ACTG'[(w >> (2 * t)) & 3] for t in range(W))


def classify(kmers):
    """special = self-overlapping (re.finditer's greedy rule matters) or longer than the window."""
    return [is_bordered(k) or len(k) > W for k in kmers]


def redundant(kmers):
    """special k-mers whose every base is covered by plain k-mers of the same list occurring inside them: they never change
    the UNION coverage (tg_cov.cu drops them from the coverage tables)."""
    sp = classify(kmers)
    out = []
    for k, x in zip(kmers, sp):
        cov = [False] * len(k)
        if x:
            for q, y in zip(kmers, sp):
                if not y:
                    o = k.find(q)
                    while o >= 0:
                        cov[o:o + len(q)] = [True] * len(q)
                        o = k.find(q, o + 1)
        out.append(x and all(cov))
    return out


E_S1A, E_T, E_S2A, E_S1B, E_S2B = 1 << 7, 1 << 8, 1 << 15, 1 << 23, 1 << 24      # tg_cov.cu entry bits


def cov_entry(s, n, list_a, list_b, meta):
    """table entry of a window holding the n bases `s` (n = W: main table; n < W: an N follows, only what fits can match):
    per list the coverage mask of the longest plain k-mer that is a prefix, S1 = exact occurrence of a window-sized
    self-overlapping k-mer, S2 = the first W bases of a long k-mer, T = bases W+1.. of a long k-mer of either list."""
    e = 0
    for g, lst in enumerate((list_a, list_b)):
        spc, red = meta[g]
        for k, x, r in zip(lst, spc, red):
            if r:
                continue
            if len(k) > W:
                rem = k[W:]
                if not (n < W and len(rem) > n) and s.startswith(rem[:W]):
                    e |= E_T
            if not s.startswith(k[:W]) or (n < W and len(k) > n):
                continue
            if not x:
                e |= ((1 << len(k)) - 1) << (16 * g)
            elif len(k) > W:
                e |= E_S2B if g else E_S2A
            else:
                e |= E_S1B if g else E_S1A
    return e


def build_cov_table(list_a, list_b):
    meta = [(classify(lst), redundant(lst)) for lst in (list_a, list_b)]
    return [cov_entry(window_str(w), W, list_a, list_b, meta) for w in range(1 << (2 * W))]


def build_cov_short_tables(list_a, list_b):
    """tab_short: the entries of windows cut short by an N after d = 1 .. 6 bases, concatenated at offsets (4^d - 4) / 3
    (tg_cov.cu: COV_SHORT_OFF)."""
    meta = [(classify(lst), redundant(lst)) for lst in (list_a, list_b)]
    return [cov_entry(window_str(w)[:d], d, list_a, list_b, meta) for d in range(1, W) for w in range(1 << (2 * d))]


def kmer_at(seq, p, k):
    return p >= 0 and p + len(k) <= len(seq) and seq[p:p + len(k)] == k


def greedy(cands, blen):
    out, blocked = [], -1
    for p in cands:
        if p >= blocked:
            out.append(p)
            blocked = p + blen
    return out


def coverage_emulate(read, kmers, tab, group):
    """coverage bit list of `read` for one list of a pair table (what scan_cov_kernel leaves in its coverage plane)."""
    L = len(read)
    isn = [c not in CODE for c in read] + [True] * (W + 40)
    cov = [0] * (L + 64)
    red = redundant(kmers)
    kmers = [k for k, r in zip(kmers, red) if not r]
    sp = classify(kmers)
    cand = {k: [] for k, x in zip(kmers, sp) if x and is_bordered(k)}
    for p in range(L):
        if isn[p]:
            continue
        if not any(isn[p:p + W]):                               # clean window: table lookup
            e = tab[window_code(read, p)]
            m = (e >> (16 * group)) & 0x7f
            for t in range(W):
                if (m >> t) & 1:
                    cov[p + t] = 1
            todo = []
            if e & (E_S1B if group else E_S1A):
                todo += [k for k, x in zip(kmers, sp) if x and len(k) <= W]
            if e & (E_S2B if group else E_S2A):
                todo += [k for k, x in zip(kmers, sp) if x and len(k) > W]
        else:                                                   # an N within reach: every k-mer is tested directly
            todo = kmers
        for k in todo:
            if kmer_at(read, p, k) and not any(isn[p:p + len(k)]):
                if k in cand:
                    cand[k].append(p)
                else:
                    for t in range(len(k)):
                        cov[p + t] = 1
    for k, c in cand.items():
        for p in greedy(c, len(k)):
            for t in range(len(k)):
                cov[p + t] = 1
    return cov[:L]


def density_from_cov(cov, w):
    L = len(cov)
    pre = [0]
    for c in cov:
        pre.append(pre[-1] + c)
    return [pre[i + w + 1] - pre[i + 1] for i in range(max(L - w, 0))]


# ---- hits -------------------------------------------------------------------------------------------------------
def build_hit_table(kmers):
    """tab[w] = set id | flag << 15; sets[id] = bit set of the plain k-mers that are a prefix of the window."""
    sp = classify(kmers)
    sets, ids, tab = [0], {0: 0}, [0] * (1 << (2 * W))
    for w in range(1 << (2 * W)):
        s = window_str(w)
        m = 0
        for k, (km, x) in enumerate(zip(kmers, sp)):
            if not x and s.startswith(km):
                m |= 1 << k
        if m not in ids:
            ids[m] = len(sets)
            sets.append(m)
        f = any(x and s.startswith(km[:W]) for km, x in zip(kmers, sp))
        tab[w] = ids[m] | (int(f) << 15)
    return tab, sets


def hits_emulate(seq, kmers, issub, tab, sets):
    """get_nonoverlapping_kmer_hits as the set algebra hits_kernel runs (SURVEY 9-E closed form, event form):
    R -> C (cover sets) -> Surv -> P / Start / ChainEnd -> one sweep over block-start segments."""
    K, L = len(kmers), len(seq)
    lens = [len(k) for k in kmers]
    maxlen = max(lens)
    sp = classify(kmers)
    isn = [c not in CODE for c in seq] + [True] * (maxlen + W + 8)
    sup = [sum(1 << j for j in issub[k]) for k in range(K)]
    # R: selected raw matches per position
    R = [0] * (L + maxlen + 2)
    cand = {k: [] for k in range(K) if sp[k] and is_bordered(kmers[k])}
    for p in range(L):
        if isn[p]:
            continue
        if not any(isn[p:p + W]):
            e = tab[window_code(seq, p)]
            R[p] |= sets[e & 0x7fff]
            todo = [k for k in range(K) if sp[k]] if e >> 15 else []
        else:
            todo = range(K)
        for k in todo:
            if kmer_at(seq, p, kmers[k]) and not any(isn[p:p + lens[k]]):
                if k in cand:
                    cand[k].append(p)
                else:
                    R[p] |= 1 << k
    for k, c in cand.items():
        for p in greedy(c, lens[k]):
            R[p] |= 1 << k
    lengt = [sum(1 << k for k in range(K) if lens[k] > d) for d in range(maxlen)]
    by_len = {}
    for k in range(K):
        by_len[lens[k]] = by_len.get(lens[k], 0) | (1 << k)
    # C(p) = k-mers whose selected match covers position p
    Cs = [0] * (L + maxlen + 2)
    for p in range(L + maxlen):
        c = 0
        for d in range(maxlen):
            if p - d >= 0:
                c |= R[p - d] & lengt[d]
        Cs[p] = c
    # Surv: k at p dies if a superstring k-mer covers any of [p, p + len_k)
    S = [0] * (L + maxlen + 2)
    for p in range(L):
        r = R[p]
        if not r:
            continue
        D, acc = {}, 0
        for d in range(maxlen):
            acc |= Cs[p + d]
            D[d + 1] = acc
        s = 0
        for k in range(K):
            if (r >> k) & 1 and not (sup[k] & D[lens[k]]):
                s |= 1 << k
        S[p] = s
    # P(x) = k-mers with a surviving match ending exactly at x; block starts and chain ends
    out = [[] for _ in range(K)]
    open_set, s_open = 0, -1
    for x in range(L + 1):
        P = 0
        for ln, m in by_len.items():
            if x - ln >= 0:
                P |= S[x - ln] & m
        start = S[x] & ~P
        end = P & ~S[x]
        if start:
            for k in range(K):
                if (open_set >> k) & 1:
                    out[k].append([s_open, x])
            open_set, s_open = start, x
        elif end & open_set:
            for k in range(K):
                if ((end & open_set) >> k) & 1:
                    out[k].append([s_open, x])
            open_set &= ~end
    assert open_set == 0
    return out
