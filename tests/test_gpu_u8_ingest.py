"""8-bit unsigned IQ ingest (ADSB_FMT_CU8, the RTL-SDR wire format; SURVEY.md 8f-4 "int8/int16 IQ widening on
device"): parity of the fused path with the oracle on the widened samples."""
import numpy as np
import pytest

import gr_adsb_b200 as g
from gr_adsb_b200 import synth
from oracle import c_oracle as co
from test_gpu_parity import check_fused, ctx_for

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sps,seed", [(2, 41), (8, 42), (6, 43), (4, 44)])
def test_cu8_input_matches_oracle(sps, seed):
    """Interleaved uint8 IQ (RTL-SDR wire format) widened on the device, value = (u - 127.5) / 128 (exact in
    float32), must give exactly what the complex64 path gives on the widened samples -- and the |IQ|^2 output
    stream of the fused block must be the oracle's mag2 bit for bit."""
    n = (1 << 20) * (2 if sps >= 6 else 1)
    d = synth.synth_numpy(n, sps=sps, n_bursts=int(n * 2e-3 * 2 / sps), seed=seed, noise_var=1e-3, amp_db=(-20.0, 0.0))
    q = np.clip(np.round(d["iq"].view(np.float32).reshape(-1, 2) * 128.0 + 127.5), 0, 255).astype(np.uint8)
    iq = ((q.astype(np.float32) - np.float32(127.5)) * np.float32(2.0 ** -7)).view(np.complex64).reshape(-1)
    x = co.mag2(iq)
    ref = co.run_stream(x, sps * 1e6, 0.01, None)
    ctx = ctx_for(sps, n)
    hits = ctx.burst_work(q, threshold=0.01)
    crc = np.array([co.crc_pass(b) for b in ref["pdu_bits"]], np.uint8)
    check_fused(hits, sps, ref["tag_off"], ref["tag_snr"], ref["pdu_off"], ref["pdu_bits"], crc)
    assert len(hits) > 20
    out = np.empty(1 << 16, np.float32)
    ctx.burst_mag_out(out)
    assert np.array_equal(out.view(np.uint32), x[-len(out):].view(np.uint32))
