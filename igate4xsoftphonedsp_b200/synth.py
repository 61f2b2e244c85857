"""Seeded synthetic inputs of SURVEY.md section 8(d) / BASELINE.md.

Pure data generation (numpy on the host, torch on the device for the large
bench shapes); no codec arithmetic happens here -- PCM is turned into G.711
codes by whoever calls this (tests: the oracle; bench: the GPU encoder).

  tone        x[n] = round(A sin(2 pi 1000 n / 8000)), A = 16384          (cfg1)
  noise+tone  per channel c: f = 300 + 100*(c % 32) Hz,
              A in {1000, 4000, 16000}[(c // 32) % 3], plus uniform integer
              noise in [-512, 512] from a counter hash seeded 0xED137 + c,
              clamped to int16                                        (cfg2-5)
  law         A-law if c % 2 == 0 else u-law
  gates       leg g of a bridge is open in frame f iff ((f // 25) + g) % 4 < 2;
              gain_q7 = 256 (the reference's 2.0) when open, 0 when shut
  ED-137 ctl  ptt = gate, pttpriority = 1 + g % 4, pttid = 0, sql = !ptt,
              bssi = (b + g) % 32
"""
import numpy as np

FRAME = 160
SEED0 = 0xED137
_M64 = (1 << 64) - 1
_K1, _K2, _K3 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB


def tone_1k(nsamples, amp=16384):
    n = np.arange(nsamples, dtype=np.float64)
    return np.round(amp * np.sin(2 * np.pi * 1000.0 * n / 8000.0)).astype(np.int16)


def _noise_np(ch, n):
    """integer noise in [-512, 512]; ch [C,1] uint64, n [1,N] uint64 (wrapping)."""
    with np.errstate(over="ignore"):
        h = (SEED0 + ch) * np.uint64(_K1) + n * np.uint64(_K2)
        h ^= h >> np.uint64(31)
        h *= np.uint64(_K3)
        h ^= h >> np.uint64(29)
    return ((h >> np.uint64(33)) % np.uint64(1025)).astype(np.int64) - 512


def channel_params(ch):
    ch = np.asarray(ch)
    freq = 300.0 + 100.0 * (ch % 32)
    amp = np.array([1000.0, 4000.0, 16000.0])[(ch // 32) % 3]
    return freq, amp


def pcm_noise_tone(F, C, ch0=0, f0=0):
    """int16 PCM [F][C][160] for channels ch0..ch0+C, frames f0..f0+F."""
    ch = np.arange(ch0, ch0 + C, dtype=np.uint64).reshape(C, 1)
    n = np.arange(f0 * FRAME, (f0 + F) * FRAME, dtype=np.uint64).reshape(1, F * FRAME)
    freq, amp = channel_params(ch.astype(np.int64))
    t = n.astype(np.float64) / 8000.0
    x = np.round(amp * np.sin(2 * np.pi * freq * t)).astype(np.int64) + _noise_np(ch, n)
    x = np.clip(x, -32768, 32767).astype(np.int16)          # [C][F*160]
    return np.ascontiguousarray(x.reshape(C, F, FRAME).transpose(1, 0, 2))


def laws(C, ch0=0):
    return ((np.arange(ch0, ch0 + C) % 2) != 0).astype(np.uint8)   # even: A-law(0), odd: u-law(1)


def out_laws(B, b0=0):
    return (np.arange(b0, b0 + B) % 2).astype(np.uint8)


def gates(F, B, G, f0=0):
    f = np.arange(f0, f0 + F).reshape(F, 1, 1)
    g = np.arange(G).reshape(1, 1, G)
    open_ = (((f // 25) + g) % 4) < 2
    return np.broadcast_to(open_, (F, B, G)).reshape(F, B * G)


def gains(F, B, G, f0=0, open_q7=256):
    return (gates(F, B, G, f0).astype(np.uint16) * np.uint16(open_q7)).astype(np.uint16)


def ed137_ctl(F, B, G, ctl_dtype, b0=0, f0=0):
    gate = gates(F, B, G, f0).reshape(F, B, G)
    ctl = np.zeros((F, B, G), dtype=ctl_dtype)
    g = np.arange(G).reshape(1, 1, G)
    b = np.arange(b0, b0 + B).reshape(1, B, 1)
    ctl["pttstatus"] = gate
    ctl["sqlstatus"] = ~gate
    ctl["pttpriority"] = np.broadcast_to(1 + g % 4, (F, B, G))
    ctl["ed137_bssi"] = np.broadcast_to((b + g) % 32, (F, B, G))
    return ctl.reshape(F, B * G)


def rtp12(F, C, pt, ch0=0, f0=0):
    """PJSIP-style 12-byte RTP headers [F][C][12]: V=2, seq=f, ts=160 f, ssrc=channel."""
    h = np.zeros((F, C, 12), dtype=np.uint8)
    f = np.arange(f0, f0 + F, dtype=np.uint32).reshape(F, 1)
    c = np.arange(ch0, ch0 + C, dtype=np.uint32).reshape(1, C)
    h[..., 0] = 0x80
    h[..., 1] = np.broadcast_to(np.asarray(pt, dtype=np.uint8).reshape(1, -1), (F, C))
    seq = np.broadcast_to(f & 0xFFFF, (F, C))
    ts = np.broadcast_to(f * 160, (F, C))
    ssrc = np.broadcast_to(c, (F, C))
    h[..., 2] = seq >> 8
    h[..., 3] = seq & 0xFF
    for k in range(4):
        h[..., 4 + k] = (ts >> (24 - 8 * k)) & 0xFF
        h[..., 8 + k] = (ssrc >> (24 - 8 * k)) & 0xFF
    return h


# ---------------------------------------------------------------- device side
def pcm_noise_tone_torch(F, C, device, ch0=0, f0=0, frame_chunk=64):
    """Same signal family generated on the GPU (bench-size inputs), int16 [F][C][160].
    Plumbing only: torch is used for RNG-free data generation, not for the hot path."""
    import torch
    out = torch.empty((F, C, FRAME), dtype=torch.int16, device=device)
    ch = torch.arange(ch0, ch0 + C, dtype=torch.int64, device=device).view(1, C, 1)
    freq = 300.0 + 100.0 * (ch % 32).double()
    amp = torch.tensor([1000.0, 4000.0, 16000.0], dtype=torch.float64, device=device)[(ch // 32) % 3]

    def s64(v):   # python int -> wrapped signed 64-bit
        v &= _M64
        return v - (1 << 64) if v >= (1 << 63) else v

    for a in range(0, F, frame_chunk):
        nf = min(frame_chunk, F - a)
        n = (torch.arange((f0 + a) * FRAME, (f0 + a + nf) * FRAME, dtype=torch.int64, device=device)
             .view(nf, 1, FRAME))
        h = (SEED0 + ch) * s64(_K1) + n * s64(_K2)
        h = h ^ ((h >> 31) & ((1 << 33) - 1))
        h = h * s64(_K3)
        h = h ^ ((h >> 29) & ((1 << 35) - 1))
        noise = ((h >> 33) & ((1 << 31) - 1)) % 1025 - 512
        x = torch.round(amp * torch.sin(2 * torch.pi * freq * (n.double() / 8000.0))).long() + noise
        out[a:a + nf] = x.clamp_(-32768, 32767).to(torch.int16)
    return out
