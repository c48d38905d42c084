"""Independent Python model of the teekay/ALAC.NET frame decoder.

TEST INFRASTRUCTURE ONLY -- the second, independent restatement used to pin
the C oracle (oracle/alac_oracle.c).  It was written from a direct reading of
the reference source, NOT from the C oracle, and deliberately uses a different
formulation everywhere it can (a pure bit-position cursor instead of the
byte-index/accumulator pair, arbitrary-precision Python ints wrapped on demand
instead of uint32 arithmetic, per-function closures instead of structs) so
that a shared transcription slip is unlikely.  Pure-Python loops: use it on
small frames only.

Citations are relative to /root/reference/ALACDecoder/.
Only VALID streams are modelled; where the reference would throw or read
stale scratch (bad tag, prediction type != 0, N beyond its buffers) the model
raises `Unmodelled` -- the status-code policy for those lives in the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass

RICE_THRESHOLD = 8        # AlacFile.cs:61


class Unmodelled(Exception):
    pass


# ---- C# int helpers ---------------------------------------
def i32(v: int) -> int:
    """wrap to a C# `int`"""
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def shl(v: int, n: int) -> int:
    return i32(v << (n & 31))


def sar(v: int, n: int) -> int:
    return i32(v) >> (n & 31)          # Python >> on a negative int is arithmetic


def sext(v: int, bits: int) -> int:
    """(v << (32-bits)) >> (32-bits)  -- AlacFile.cs:278-279, :289-290, :309-310"""
    mv = 32 - bits
    return sar(shl(v, mv), mv)


def tdiv(a: int, b: int) -> int:
    """C# integer division truncates toward zero"""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def clz_quirk(v: int) -> int:
    """CountLeadingZeros (AlacFile.cs:170-191): byte-wise scan of a 32-bit int
    that falls off the end for 0 and returns 32 + 8."""
    v = i32(v)
    if v == 0:
        return 40
    if v < 0:
        return 0
    return 32 - v.bit_length()


@dataclass
class Cookie:
    """What AlacFile.SetInfo keeps (AlacFile.cs:63-93) + the container channel
    count handed to the constructor (AlacContext.cs:54)."""
    sample_size: int = 16
    num_channels: int = 2
    max_samples_per_frame: int = 4096
    rice_history_mult: int = 40
    rice_initial_history: int = 10
    rice_kmodifier: int = 14

    @staticmethod
    def from_codec_data(cd: bytes, num_channels: int, sample_size: int) -> "Cookie":
        """cd = DemuxResT.CodecData as bytes; offsets as in SetInfo (AlacFile.cs:63-93)."""
        msf = int.from_bytes(cd[24:28], "big")
        return Cookie(sample_size, num_channels, msf, cd[30], cd[31], cd[32])


class Bits:
    """MSB-first cursor over a frame.  Readbits16 (AlacFile.cs:101-118) takes the
    top `bits` bits of a 24-bit look-ahead shifted by the accumulator, which is
    the same as reading `bits` bits at the absolute bit position; Readbits
    (:125-129) glues two such reads high-half first; Unreadbits (:145-152) steps
    the position back.  Bytes past the frame read as 0."""

    def __init__(self, data: bytes):
        self.data = data
        self.pos = 0
        self.value = int.from_bytes(data, "big") if data else 0
        self.nbits = len(data) * 8

    def read(self, n: int) -> int:
        if n == 0:
            return 0
        end = self.pos + n
        if end <= self.nbits:
            v = (self.value >> (self.nbits - end)) & ((1 << n) - 1)
        else:   # zero fill past the frame
            have = max(0, self.nbits - self.pos)
            v = (self.value & ((1 << have) - 1)) << (n - have) if have else 0
        self.pos = end
        return v

    def unread(self, n: int) -> None:
        self.pos -= n


def decode_symbol(br: Bits, raw_bits: int, k: int, mask: int) -> int:
    """EntropyDecodeValue (AlacFile.cs:193-212)."""
    x = 0
    while x <= RICE_THRESHOLD and br.read(1) != 0:      # up to nine 1-bits, the 0 is consumed
        x += 1
    if x > RICE_THRESHOLD:
        return i32(br.read(raw_bits)) & i32(0xFFFFFFFF >> ((32 - raw_bits) & 31))
    if k == 1:
        return x
    extra = br.read(k)
    x = i32(x * (i32((1 << (k & 31)) - 1) & i32(mask)))
    if extra > 1:
        x = i32(x + extra - 1)
    else:
        br.unread(1)
    return x


def rice_decode(br: Bits, n: int, rss: int, initial_history: int, kmod: int, mult: int, mask: int) -> list[int]:
    """EntropyRiceDecode (AlacFile.cs:214-252)."""
    out = [0] * (n + 0x20000)       # runs may write past n (the reference has a 16384 scratch)
    history = initial_history
    sign_mod = 0
    count = 0
    while count < n:
        t = 31 - kmod - clz_quirk(sar(history, 9) + 3)
        k = t + kmod if t < 0 else kmod
        dv = i32(decode_symbol(br, rss, k, 0xFFFFFFFF) + sign_mod)
        half = tdiv(i32(dv + 1), 2)
        out[count] = -half if dv & 1 else half
        sign_mod = 0
        if dv > 0xFFFF:
            history = 0xFFFF
        else:
            history = i32(history + i32(dv * mult) - sar(i32(history * mult), 9))
        if history < 128 and count + 1 < n:
            if history < 0:
                raise Unmodelled("negative rice history")
            sign_mod = 1
            k = clz_quirk(history) + tdiv(history + 16, 64) - 24
            run = decode_symbol(br, 16, k, mask)
            if run > 0:
                if count + 1 + run > len(out):
                    raise Unmodelled("zero run beyond the scratch buffer")
                # zeros are already there
                count += run
            if run > 0xFFFF:
                sign_mod = 0
            history = 0
        count += 1
    return out[:n]


def predict(e: list[int], n: int, rss: int, coef: list[int], order: int, quant: int) -> list[int]:
    """PredictorDecompressFirAdapt (AlacFile.cs:256-336); returns the output
    samples and updates `coef` in place, as the reference does."""
    o = list(e)
    if order == 0:
        if n > 4096:
            raise Unmodelled("order 0 with more samples than the Array.Copy length allows")
        return o
    if n <= 1:
        return o
    if order == 31:
        for i in range(n - 1):
            o[i + 1] = sext(i32(o[i] + e[i + 1]), rss)
        return o
    for i in range(order):
        if i + 1 >= n:      # the reference indexes its 16384 scratch here; stay within n
            break
        o[i + 1] = sext(i32(o[i] + e[i + 1]), rss)
    base = 0
    for i in range(order + 1, n):
        err = e[i]
        s = 0
        for j in range(order):
            s = i32(s + i32(i32(o[base + order - j] - o[base]) * coef[j]))
        v = i32(shl(1, quant - 1) + s)
        v = sar(v, quant)
        v = i32(i32(v + o[base]) + err)
        o[base + order + 1] = sext(v, rss)
        if err != 0:
            positive = err > 0
            p = order - 1
            while p >= 0 and ((err > 0) if positive else (err < 0)):
                val = i32(o[base] - o[base + order - p])
                sg = -1 if val < 0 else (1 if val > 0 else 0)
                if not positive:
                    sg = -sg
                coef[p] = i32(coef[p] - sg)
                val = i32(val * sg)
                err = i32(err - i32(sar(val, quant) * (order - p)))
                p -= 1
        base += 1
    return o


def decode_frame_ints(ck: Cookie, frame: bytes) -> tuple[list[int], int]:
    """AlacFile.DecodeFrame (AlacFile.cs:428-719) -> (outbuffer ints, outputsize bytes).
    16-bit: one int per sample; 24-bit: one byte-valued int per output byte."""
    br = Bits(frame)
    ss, nch = ck.sample_size, ck.num_channels
    if ss not in (16, 24):
        raise Unmodelled("sample size")
    bps = (ss // 8) * nch                       # AlacFile.cs:19
    n = ck.max_samples_per_frame
    tag = br.read(3)
    if tag > 1:
        raise Unmodelled("element tag")
    stereo = tag == 1
    br.read(4)
    br.read(12)
    hassize = br.read(1)
    ub = br.read(2)
    escape = br.read(1)
    if hassize:
        n = i32(br.read(32))
        if n < 0 or n > 16384 or n * bps > 65536:
            raise Unmodelled("sample count beyond the reference's buffers")
    outputsize = n * bps
    rss = ss - ub * 8 + (1 if stereo else 0)
    ech = 2 if stereo else 1
    chans: list[list[int]] = []
    shifts: list[list[int]] = [[], []]
    mix_shift = mix_weight = 0
    if not escape:
        if rss < 1:
            raise Unmodelled("read sample size")
        a, b = br.read(8), br.read(8)
        if stereo:
            mix_shift, mix_weight = a, b
        hdr = []
        for _ in range(ech):
            ptype, quant, rmod, order = br.read(4), br.read(4), br.read(3), br.read(5)
            coef = []
            for _ in range(order):
                c = br.read(16)
                coef.append(c - 65536 if c > 32767 else c)
            if ptype != 0:
                raise Unmodelled("prediction type")
            hdr.append((quant, rmod, order, coef))
        if ub:
            for _ in range(n):
                for c in range(ech):
                    shifts[c].append(br.read(ub * 8))
        mask = i32((1 << ck.rice_kmodifier) - 1)
        for quant, rmod, order, coef in hdr:
            e = rice_decode(br, n, rss, ck.rice_initial_history, ck.rice_kmodifier,
                            rmod * tdiv(ck.rice_history_mult, 4), mask)
            chans.append(predict(e, n, rss, coef, order, quant))
    else:
        for c in range(ech):
            chans.append([0] * n)
        for i in range(n):
            for c in range(ech):
                if ss <= 16:
                    chans[c][i] = sext(br.read(ss), ss)
                else:
                    v = shl(br.read(16), ss - 16) | br.read(ss - 16)
                    x = v & 0xFFFFFF
                    chans[c][i] = (x ^ 0x800000) - 0x800000
        ub = 0
    if br.pos > len(frame) * 8:
        raise Unmodelled("bitstream overrun")

    # un-mix (Deinterlace16/24, AlacFile.cs:338-421) or mono (AlacFile.cs:527-575)
    left, right = chans[0], ([0] * n)
    if stereo:
        bch = chans[1]
        if mix_weight != 0:
            right = [i32(chans[0][i] - sar(i32(bch[i] * mix_weight), mix_shift)) for i in range(n)]
            left = [i32(right[i] + bch[i]) for i in range(n)]
        else:
            left, right = chans[0], bch
    if ss == 24 and ub:
        m = i32(~(0xFFFFFFFF << (ub * 8)))
        left = [shl(left[i], ub * 8) | (shifts[0][i] & m) for i in range(n)]
        if stereo:
            right = [shl(right[i], ub * 8) | (shifts[1][i] & m) for i in range(n)]
    # the reference writes slot i*nch (+1) and lets iteration i+1 overwrite what a
    # 1-channel container does not own: net effect = `nch` channels, left first
    out: list[int] = []
    for i in range(n):
        vals = (left[i], right[i])[:nch]
        for v in vals:
            if ss == 16:
                out.append(v)
            else:
                out.extend((v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF))
    return out, outputsize


def read_frame(ck: Cookie, frame: bytes) -> bytes:
    """AlacContext.Read for one frame: DecodeFrame + FormatSamples (AlacContext.cs:163-172, :214-256)."""
    ints, size = decode_frame_ints(ck, frame)
    dst = bytearray()
    if ck.sample_size == 16:                      # bps 2: low byte, then (uint)temp >> 8
        for v in ints[: size // 2]:
            dst.append(v & 0xFF)
            dst.append((v >> 8) & 0xFF)
    else:                                         # bps 3: one byte per int
        for v in ints[:size]:
            dst.append(v & 0xFF)
    return bytes(dst)


def decode_track(ck: Cookie, mdat: bytes, stsz) -> bytes:
    """Sequential frame pump (AlacContext.cs:179-204): frame i = next stsz[i] bytes."""
    pos = 0
    out = bytearray()
    for sz in stsz:
        sz = int(sz)
        out += read_frame(ck, mdat[pos:pos + sz])
        pos += sz
    return bytes(out)
