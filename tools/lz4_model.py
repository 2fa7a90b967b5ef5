"""Executable model of the device LZ4 decoder (snappydata_b200/csrc/sd_lz4.cu), used to check its ring / flush /
dependency-round logic on the host: the same group parse, per-lane copies in rounds, shared-memory output ring with
"shifted" positions, 16-byte flushes and near (ring) / far (HBM) source reads -- with assertions where the kernel
relies on an invariant (a ring slot still holds the byte it is read for; a far byte has been flushed; a match
never reads a byte that a pending sequence has yet to produce).  The constants are read from the .cu file, so
changing a shape (LzDefault / LzDense) or LZ_MAX_* there and running this (or tests/test_lz4_model.py) re-checks the invariants.

    python tools/lz4_model.py          # decodes a set of adversarial blocks at several output alignments
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snappydata_b200.column_format import compress_lz4  # noqa: E402


def kernel_constants(which=None):
    """Constants of one kernel shape: `typedef LzCfg<WARPS, WIN, IN, PIECE> <which>;` + the shared LZ_MAX_*."""
    which = which or os.environ.get("LZ4_MODEL_CFG", "LzDefault")
    text = open(os.path.join(ROOT, "snappydata_b200", "csrc", "sd_lz4.cu")).read()
    m = re.search(r"typedef LzCfg<(\d+), (\d+), (\d+), (\d+)> %s;" % which, text)
    out = {"LZ_WIN": int(m.group(2)), "LZ_IN": int(m.group(3)), "LZ_PIECE": int(m.group(4))}
    for name in ("LZ_MAX_LIT", "LZ_MAX_ML"):
        out[name] = int(re.search(r"constexpr int %s = (\d+);" % name, text).group(1))
    return out


def use(which):
    """Switch the model to another kernel shape (LzDefault | LzDense)."""
    global K, WIN, MAX_LIT, MAX_ML, PIECE, M
    K = kernel_constants(which)
    WIN, MAX_LIT, MAX_ML, PIECE = K["LZ_WIN"], K["LZ_MAX_LIT"], K["LZ_MAX_ML"], K["LZ_PIECE"]
    M = WIN - 1


K = kernel_constants()
WIN, MAX_LIT, MAX_ML, PIECE = K["LZ_WIN"], K["LZ_MAX_LIT"], K["LZ_MAX_ML"], K["LZ_PIECE"]
M = WIN - 1

WP = 256          # input bytes examined per step by the window parse (LZ_WP in the kernel)
STOP = 0xFFFF     # "the sequence starting here needs the checked path" (long run, > 1 length byte)


def window_parse(src, s, o, wofs):
    """Model of the kernel's window parse (sd_lz4.cu, PARSE == 1): every position of the window is parsed speculatively as
    if a token started there (independent work, no chain), two doubling passes turn `next token` into a 4-step jump, the
    chain from position 0 is then walked 4 sequences at a time (8 dependent steps for 32 sequences) and the lanes fill in
    the members in between; fields are extracted per lane and output positions come from a prefix sum.
    Returns (records, next s, next o); no record when the first sequence needs the checked path."""
    def spec(p):
        tok = src[s + p]; lit = tok >> 4; ml = tok & 15; q = p + 1
        if lit == 15:
            e = src[s + q]; q += 1; lit += e
            if e == 255: return STOP, None
        if lit > MAX_LIT: return STOP, None
        lit_src = q; q += lit
        q_off = q; q += 2
        if ml == 15:
            e = src[s + q]; q += 1; ml += e
            if e == 255: return STOP, None
        ml += 4
        if ml > MAX_ML: return STOP, None
        return q, (lit_src, lit, q_off, ml)
    nxt = [spec(p)[0] for p in range(WP)]
    def hop(table, p):          # member after p through `table`, or STOP when there is none inside the window
        return table[p] if p < WP else STOP
    j1 = [STOP if nxt[p] >= WP else nxt[nxt[p]] for p in range(WP)]
    j1 = [STOP if v >= WP else v for v in j1]            # only in-window positions are members
    j2 = [STOP if j1[p] >= WP else j1[j1[p]] for p in range(WP)]
    j2 = [STOP if v >= WP else v for v in j2]
    anchors = []
    a = 0
    for i in range(8):
        anchors.append(a)
        a = j2[a] if a < WP else STOP
    recs, last_next = [], None
    lane_pos = []
    for L in range(32):
        a = anchors[L >> 2]; r = L & 3
        if a >= WP: pos = STOP
        elif r == 0: pos = a
        elif r == 1: pos = nxt[a] if nxt[a] < WP else STOP
        elif r == 2: pos = j1[a]
        else: pos = (nxt[j1[a]] if j1[a] < WP and nxt[j1[a]] < WP else STOP)
        lane_pos.append(pos)
    valid = [pos < WP and nxt[pos] != STOP for pos in lane_pos]
    n = 0
    while n < 32 and valid[n]: n += 1
    assert not any(valid[n:]) or True   # (lanes after the first invalid one are ignored, as in the kernel)
    lens = []
    for L in range(n):
        q, (lit_src, lit, q_off, ml) = spec(lane_pos[L])
        lens.append((lit_src, lit, q_off, ml, q))
    oo = o
    for (lit_src, lit, q_off, ml, q) in lens:
        off = src[s + q_off] | (src[s + q_off + 1] << 8)
        assert off != 0 and off <= oo - wofs + lit, "bad offset"
        recs.append((s + lit_src, lit, oo + lit, off, ml))
        oo += lit + ml
    new_s = s + (lens[-1][4] if n else 0)
    return recs, new_s, oo


def decode(src, n_out, wofs, parse=0):
    """parse=0: the serial group parse; parse=1: the window parse where a whole group's input and output remain"""
    n_src=len(src); end=n_out+wofs
    dst_al=bytearray(end+64); written=bytearray(end+64)   # HBM
    win=bytearray(WIN); ring_pos=[-1]*WIN                  # which P each slot holds
    st=dict(s=0,o=wofs,flushed=wofs)
    def wr(P,v): win[P&M]=v; ring_pos[P&M]=P
    def rr(P):
        assert ring_pos[P&M]==P, f"ring slot for {P} holds {ring_pos[P&M]}"
        return win[P&M]
    def far(P):
        assert written[P], f"far read of unflushed byte {P} (flushed {st['flushed']})"
        return dst_al[P]
    def flush(upto):
        lim=upto&~15
        if lim<=st['flushed']: return
        if st['flushed']&15:
            he=(st['flushed']+15)&~15
            for P in range(st['flushed'],he): dst_al[P]=rr(P); written[P]=1
            st['flushed']=he
        for v in range(st['flushed']>>4, lim>>4):
            for u in range(16): P=v*16+u; dst_al[P]=rr(P); written[P]=1
        st['flushed']=lim
    finished=False
    while not finished:
        recs=[]; big=None
        s=st['s']; o=st['o']
        use_window = parse==1 and n_src-s >= WP+64 and end-o >= 32*(MAX_LIT+MAX_ML)
        if use_window:
            wrecs, ws, wo = window_parse(src, s, o, wofs)
        while len(recs)<32:
            if use_window and wrecs:      # the serial loop below re-derives the same records: cross-check, then take them
                pass
            if s>=n_src: finished=True; break
            tok=src[s]; s+=1; lit=tok>>4
            if lit==15:
                while True:
                    b=src[s]; s+=1; lit+=b
                    if b!=255: break
            assert lit<=n_src-s and lit<=end-o
            lit_src=s; s+=lit; last=s>=n_src; off=ml=0
            if not last:
                off=src[s]|(src[s+1]<<8); s+=2; ml=tok&15
                if ml==15:
                    while True:
                        b=src[s]; s+=1; ml+=b
                        if b!=255: break
                ml+=4
                assert off!=0 and off<=o-wofs+lit and ml<=end-o-lit
            if lit>MAX_LIT or ml>MAX_ML: big=(lit_src,lit,off,ml,last); break
            recs.append((lit_src,lit,o+lit,off,ml)); o+=lit+ml
            if last: finished=True; break
        if use_window and wrecs:
            # the window parse must return a prefix of what the serial parse finds from the same position
            assert recs[:len(wrecs)] == wrecs, (recs[:3], wrecs[:3])
            if len(wrecs) < len(recs) or big or finished:
                # take exactly the window's group; the rest is parsed again in the next step
                recs = wrecs; big = None; finished = False; s = ws; o = wo
            assert (s, o) == (ws, wo), ((s, o), (ws, wo))
        st['s']=s; st['o']=o
        n=len(recs)
        if n:
            group_end=o
            for (lit_src,lit,mdst,off,ml) in recs:
                for k in range(lit): wr(mdst-lit+k, src[lit_src+k])
            pending=[r[4]>0 for r in recs]
            while any(pending):
                k=pending.index(True); W=recs[k][2]
                writes=[]
                for i,(lit_src,lit,mdst,off,ml) in enumerate(recs):
                    if not pending[i]: continue
                    msrc=mdst-off; dep_end=msrc+min(ml,off); near=group_end-msrc<=WIN
                    if i==k or dep_end<=W:
                        own={}
                        for q in range(ml):
                            P=msrc+q
                            if P in own: v=own[P]
                            else:
                                if near:
                                    v=rr(P); assert P<W, f"near read of byte {P} that a pending sequence produces (ready_below {W})"
                                else: v=far(P)
                            own[mdst+q]=v
                        writes.append(own); pending[i]=False
                for own in writes:
                    for P,v in own.items(): wr(P,v)
            flush(o)
        if big:
            lit_src,lit,off,ml,last=big
            done=0
            while done<lit:
                piece=min(PIECE,lit-done)
                for i in range(piece): wr(o+i, src[lit_src+done+i])
                o+=piece; done+=piece; flush(o)
            m0=o; done=0
            while done<ml:
                piece=min(PIECE,ml-done); piece_end=o+piece
                vals=[]
                for i in range(piece):
                    P=m0-off+done+i if off>=ml else m0-off+(done+i)%off
                    assert P<o
                    vals.append(rr(P) if piece_end-P<=WIN else far(P))
                for i,v in enumerate(vals): wr(o+i,v)
                o+=piece; done+=piece; flush(o)
            st['o']=o
            if last: finished=True
    assert st['o']==end
    flush(st['o'])
    for P in range(st['flushed'],end): dst_al[P]=rr(P); written[P]=1
    assert all(written[wofs:end])
    return bytes(dst_al[wofs:end])


def blocks(rng, scale=1):
    a30 = rng.bytes(30_000)
    return [
        bytes(40_000 // scale),
        b"".join(bytes([i + 1]) * (i + 1) * (700 // scale) for i in range(9)),
        b"".join((b"abcdefghi"[:k] * 3000)[:9000 // scale] for k in range(1, 10)),
        rng.bytes(50_000 // scale),
        a30 + rng.bytes(25_000) + a30 + rng.bytes(100) + a30[5_000:9_000],
        rng.integers(1, 51, 20_000 // scale).astype(np.float64).tobytes(),
        rng.integers(0, 3, 30_000 // scale).astype(np.int16).tobytes(),
        rng.integers(8000, 10_500, 20_000 // scale).astype(np.int32).tobytes(),
        b"".join(rng.bytes(int(n)) + bytes(int(m)) for n, m in zip(rng.integers(0, 80, 400 // scale), rng.integers(4, 300, 400 // scale))),
        rng.bytes(13) + bytes(27),
        b"xy",
    ]


def check(scale=1, alignments=(0, 8, 5, 15), parse=0):
    rng = np.random.default_rng(5)
    for bi, b in enumerate(blocks(rng, scale)):
        env = compress_lz4(b, force=True)
        for wofs in alignments:
            assert decode(env[8:], len(b), wofs, parse) == b, (bi, wofs)
        yield bi, len(b), len(env)


if __name__ == "__main__":
    for which in ("LzDefault", "LzDense"):
        use(which)
        print(which, K)
        for parse in (0, 1):
            for bi, n, c in check(parse=parse):
                print("  ok block", bi, n, "->", c, "bytes", "(window parse)" if parse else "")
