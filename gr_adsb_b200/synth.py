"""Seeded synthetic Mode-S / ADS-B IQ streams for tests and bench.py.

Noise is circular complex Gaussian (|IQ|^2 ~ Exp(mean noise_var)); bursts are DF17
extended squitters with a valid CRC-24 parity field, PPM-modulated the way the reference
demodulator reads them (demod.py:87-95: bit 1 -> energy in the first half symbol, bit 0 ->
second half), preceded by the 8 us preamble whose half-symbol template the reference framer
matches (framer.py:50: pulses at half symbols 0, 2, 7, 9 of 16).

`synth_numpy` is used by the tests (small, bit-reproducible from the seed); `synth_torch`
builds the same kind of stream directly in HBM for the large bench buffers.
"""
import numpy as np

NUM_PREAMBLE_HALF_SYMBOLS = 16
PREAMBLE_HIGH = (0, 2, 7, 9)
NUM_BITS = 112
CRC_GENERATOR = 0x1FFF409   # 25-bit Mode S generator; decoder.py:269 holds it as a bit array


def _crc_tables():
    """T[i] = x^(111-i) mod g(x): contribution of message bit i to the 24-bit remainder."""
    t = np.zeros(NUM_BITS, dtype=np.uint32)
    r = 1
    for i in range(NUM_BITS - 1, -1, -1):
        t[i] = r
        r <<= 1
        if r & (1 << 24):
            r ^= CRC_GENERATOR
    return t


CRC_T112 = _crc_tables()


def crc24_syndrome(bits):
    """Remainder of the 112-bit (or shorter, right-aligned) message polynomial; 0 = parity OK.

    `bits` is uint8[..., L] with L <= 112, MSB (bit 0) first.
    """
    bits = np.asarray(bits, dtype=np.uint8)
    L = bits.shape[-1]
    t = CRC_T112[NUM_BITS - L:]
    return np.bitwise_xor.reduce(np.where(bits != 0, t, np.uint32(0)), axis=-1)


def crc24_of_payload(bits88):
    """24 parity bits for an 88-bit payload (what decoder.compute_crc returns, decoder.py:693-715)."""
    bits88 = np.asarray(bits88, dtype=np.uint8)
    pad = np.concatenate((bits88, np.zeros(bits88.shape[:-1] + (24,), dtype=np.uint8)), axis=-1)
    syn = crc24_syndrome(pad)
    shifts = np.arange(23, -1, -1, dtype=np.uint32)
    return ((syn[..., None] >> shifts) & 1).astype(np.uint8)


def hex_to_bits(h):
    v = int(h, 16)
    n = len(h) * 4
    return np.array([(v >> (n - 1 - i)) & 1 for i in range(n)], dtype=np.uint8)


def bits_to_hex(bits):
    v = 0
    for b in bits:
        v = (v << 1) | int(b)
    return "%0*X" % (len(bits) // 4, v)


def df17_messages(rng, n):
    """n random DF17 (CA=5) messages with valid parity -> uint8[n,112]."""
    bits = rng.integers(0, 2, size=(n, NUM_BITS), dtype=np.uint8)
    bits[:, 0:5] = (1, 0, 0, 0, 1)
    bits[:, 5:8] = (1, 0, 1)
    bits[:, 88:] = crc24_of_payload(bits[:, :88])
    return bits


def high_half_symbols(bits):
    """Half-symbol slots (relative to burst start) that carry a pulse -> int64[n,116]."""
    bits = np.asarray(bits, dtype=np.uint8).reshape(-1, NUM_BITS)
    k = np.arange(NUM_BITS, dtype=np.int64)
    data = NUM_PREAMBLE_HALF_SYMBOLS + 2 * k[None, :] + (1 - bits.astype(np.int64))
    pre = np.broadcast_to(np.array(PREAMBLE_HIGH, dtype=np.int64), (bits.shape[0], 4))
    return np.concatenate((pre, data), axis=1)


def burst_len(sps):
    return (8 + NUM_BITS) * sps


def synth_numpy(n_samples, sps=2, n_bursts=100, seed=1090, noise_var=1e-3, amp_db=(-10.0, 0.0),
                nonoverlap=False, cfo_hz=0.0, bits=None, starts=None):
    """Return dict(iq=complex64[n], starts=int64[nb], bits=uint8[nb,112], amp2=float[nb]).

    amp2 is the burst's |IQ|^2 pulse level, uniform in dB over `amp_db`.  Start indices are
    uniform over the stream; with `nonoverlap` they are drawn on a jittered grid so bursts
    (and their 63-symbol blanking interval) never collide.
    """
    assert sps >= 2 and sps % 2 == 0
    rng = np.random.default_rng(seed)
    h = sps // 2
    blen = burst_len(sps)
    noise = rng.standard_normal((n_samples, 2)).astype(np.float32) * np.float32(np.sqrt(noise_var / 2.0))
    iq = (noise[:, 0] + 1j * noise[:, 1]).astype(np.complex64)
    if bits is None:
        bits = df17_messages(rng, n_bursts)
    bits = np.asarray(bits, dtype=np.uint8).reshape(-1, NUM_BITS)
    nb = bits.shape[0]
    if starts is None:
        hi = n_samples - blen - 8 * sps
        if nonoverlap:
            slot = hi // max(nb, 1)
            assert slot >= blen + 110 + 8 * sps, "stream too short for non-overlapping bursts"
            starts = 104 + np.arange(nb, dtype=np.int64) * slot + rng.integers(0, slot - blen - 104 - 8 * sps, size=nb)
        else:
            starts = np.sort(rng.integers(104, hi, size=nb)).astype(np.int64)
    starts = np.asarray(starts, dtype=np.int64)
    amp2 = 10.0 ** (rng.uniform(amp_db[0], amp_db[1], size=nb) / 10.0)
    phase = rng.uniform(0.0, 2.0 * np.pi, size=nb)
    if nb:
        slots = high_half_symbols(bits)                                  # [nb,116]
        idx = (starts[:, None, None] + slots[:, :, None] * h + np.arange(h)[None, None, :])  # [nb,116,h]
        ph = phase[:, None, None] + 2.0 * np.pi * cfo_hz / (sps * 1e6) * (idx - starts[:, None, None])
        val = (np.sqrt(amp2)[:, None, None] * np.exp(1j * ph)).astype(np.complex64)
        np.add.at(iq, idx.ravel(), val.ravel())
    return {"iq": iq, "starts": starts, "bits": bits, "amp2": amp2}


def mag2(iq):
    """|IQ|^2 as gr-blocks complex_to_mag_squared computes it: fl32(fl32(re*re)+fl32(im*im))
    (adsb_rx.py:180; VOLK's SIMD kernels multiply then add, no FMA contraction)."""
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    v = iq.view(np.float32).reshape(-1, 2)
    return (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]).astype(np.float32)


def synth_torch(n_samples, sps, n_bursts, seed, device, noise_var=1e-3, amp_db=(-10.0, 0.0),
                start_lo=104, return_truth=True, out=None):
    """Same construction in torch, directly on `device`; returns (iq complex64[n] as float32[n,2],
    truth dict of device tensors)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    h = sps // 2
    blen = burst_len(sps)
    iq = out if out is not None else torch.empty((n_samples, 2), dtype=torch.float32, device=device)
    assert iq.shape == (n_samples, 2) and iq.is_contiguous()
    iq.normal_(0.0, float(np.sqrt(noise_var / 2.0)), generator=g)
    nb = int(n_bursts)
    if nb == 0:
        return iq, {"starts": None, "bits": None, "amp2": None}
    # messages: random payload with DF17/CA5 header, parity via the GF(2) syndrome table
    bits = torch.randint(0, 2, (nb, NUM_BITS), dtype=torch.int32, device=device, generator=g)
    bits[:, 0:5] = torch.tensor([1, 0, 0, 0, 1], dtype=torch.int32, device=device)
    bits[:, 5:8] = torch.tensor([1, 0, 1], dtype=torch.int32, device=device)
    t = torch.from_numpy(CRC_T112.astype(np.int64)).to(device)
    contrib = torch.where(bits[:, :88] != 0, t[None, :88], torch.zeros((), dtype=torch.int64, device=device))
    syn = torch.zeros(nb, dtype=torch.int64, device=device)
    for i in range(88):                       # xor-reduce (torch has no bitwise reduce)
        syn ^= contrib[:, i]
    shifts = torch.arange(23, -1, -1, device=device, dtype=torch.int64)
    bits[:, 88:] = ((syn[:, None] >> shifts[None, :]) & 1).to(torch.int32)
    hi = n_samples - blen - 8 * sps
    starts = torch.randint(start_lo, hi, (nb,), device=device, generator=g, dtype=torch.int64)
    starts, _ = torch.sort(starts)
    amp_dbv = torch.empty(nb, device=device, dtype=torch.float64).uniform_(amp_db[0], amp_db[1], generator=g)
    amp = torch.pow(10.0, amp_dbv / 20.0)
    phase = torch.empty(nb, device=device, dtype=torch.float64).uniform_(0.0, 2.0 * np.pi, generator=g)
    k = torch.arange(NUM_BITS, device=device, dtype=torch.int64)
    data = NUM_PREAMBLE_HALF_SYMBOLS + 2 * k[None, :] + (1 - bits.to(torch.int64))
    pre = torch.tensor(PREAMBLE_HIGH, device=device, dtype=torch.int64)[None, :].expand(nb, 4)
    slots = torch.cat((pre, data), dim=1)                                   # [nb,116]
    re = (amp * torch.cos(phase)).to(torch.float32)
    im = (amp * torch.sin(phase)).to(torch.float32)
    flat = iq.view(-1)
    chunk = max(1, (1 << 24) // (116 * h))
    for b0 in range(0, nb, chunk):
        b1 = min(nb, b0 + chunk)
        idx = (starts[b0:b1, None, None] + slots[b0:b1, :, None] * h
               + torch.arange(h, device=device, dtype=torch.int64)[None, None, :]).reshape(-1)
        n_per = 116 * h
        flat.index_add_(0, idx * 2, re[b0:b1].repeat_interleave(n_per))
        flat.index_add_(0, idx * 2 + 1, im[b0:b1].repeat_interleave(n_per))
    truth = {"starts": starts, "bits": bits.to(torch.uint8), "amp2": (amp * amp)} if return_truth else {}
    return iq, truth
