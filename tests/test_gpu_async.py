"""phn_recognize_async / phn_wait (two batches in flight) and the decoder stream.

The reference walks a list one file after the other (SpeechRec::ProcessFileList, srec.cpp:1246-1291), so its output cannot
depend on any overlap; here batch k+1 is enqueued while batch k is on the device (its audio goes up under batch k's nets,
batch k's decoder runs on its own stream under batch k+1's front end).  Whatever the interleaving, every batch must return
exactly what the synchronous call returns for it - in the exact mode that is the reference's output bit for bit."""
import numpy as np
import pytest

from conftest import audio_bytes, model_dir
import phnrec_b200 as pb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rec():
    r = pb.Recognizer(model_dir("PHN_CZ_SPDAT_LCRC_N1500"), device=0)
    yield r
    r.close()


def _batches(r):
    a = audio_bytes("test.raw")
    syn = r.synth_audio(16000, 12, seed=5)          # lin16: 1 s each
    b0 = [a, a[:50000], a[3000:90000]]
    b1 = [syn[i].tobytes() for i in range(12)]
    b2 = [a[:398], a[20000:], a[:402]]
    b3 = [syn[i].tobytes()[: 9000 + 700 * i] for i in range(12)]
    b4 = [a]
    return [b0, b1, b2, b3, b4]


@pytest.mark.parametrize("mode", [pb.MLP_EXACT_FP32, pb.MLP_TC_F16])
def test_pipelined_batches_equal_synchronous_calls(rec, mode):
    rec.set_mlp_mode(mode)
    batches = _batches(rec)
    want = [rec.recognize(b) for b in batches]
    got = rec.recognize_pipelined(batches)
    assert rec.pending() == 0 and len(got) == len(want)
    for wb, gb in zip(want, got):
        assert len(wb) == len(gb)
        for w, g in zip(wb, gb):
            assert np.array_equal(w.view(np.uint8), g.view(np.uint8))
    # and once more in the reverse order (slots and streams reused)
    got = rec.recognize_pipelined(batches[::-1])[::-1]
    for wb, gb in zip(want, got):
        for w, g in zip(wb, gb):
            assert np.array_equal(w.view(np.uint8), g.view(np.uint8))


def test_async_protocol_errors(rec):
    rec.set_mlp_mode(pb.MLP_TC_F16)
    a = np.frombuffer(audio_bytes("test.raw"), dtype=np.uint8)
    boff = np.array([0, a.size], dtype=np.int64)
    labels = np.zeros(2000, dtype=pb.LABEL_DTYPE)
    loff = np.zeros(2, dtype=np.int64)
    with pytest.raises(pb.PhnRecError):
        rec.wait_raw(labels, loff)                      # nothing in flight
    rec.recognize_async_raw(a.ctypes.data, boff)
    rec.recognize_async_raw(a.ctypes.data, boff)
    assert rec.pending() == 2
    with pytest.raises(pb.PhnRecError):
        rec.recognize_async_raw(a.ctypes.data, boff)    # a third one needs a wait first
    with pytest.raises(pb.PhnRecError):
        rec.recognize([a.tobytes()])                    # the synchronous call refuses to jump the queue
    small = np.zeros(1, dtype=pb.LABEL_DTYPE)
    with pytest.raises(pb.PhnRecError):
        rec.wait_raw(small, loff)                       # capacity error: the batch stays queued, loff holds the count
    assert rec.pending() == 2 and loff[1] > 1
    n0 = rec.wait_raw(labels, loff)
    first = labels[:n0].copy()
    n1 = rec.wait_raw(labels, loff)
    assert rec.pending() == 0 and n0 == n1
    assert np.array_equal(first.view(np.uint8), labels[:n1].view(np.uint8))
    sync = rec.recognize([a.tobytes()])[0]
    assert np.array_equal(sync.view(np.uint8), first.view(np.uint8))


def test_device_calls_back_to_back_keep_the_last_result(rec):
    """phn_recognize_device twice without a fetch in between (what bench.py's device loop does): the decoder of the first
    call runs on the decoder stream while the second call's front end starts; the fetch returns the second call's labels."""
    rec.set_mlp_mode(pb.MLP_TC_F16)
    rec.set_wave_format("alaw")
    try:
        n = 64
        x = rec.synth_audio(40000, n, seed=21)
        y = rec.synth_audio(40000, n, seed=22)
        want = rec.recognize([y[i].tobytes() for i in range(n)])
        boff = np.arange(n + 1, dtype=np.int64) * 40000
        dx, dy = rec.device_alloc(x.size), rec.device_alloc(y.size)
        rec.memcpy_h2d(dx, x.ctypes.data, x.size)
        rec.memcpy_h2d(dy, y.ctypes.data, y.size)
        for _ in range(3):
            rec.recognize_device(dx, boff)
            rec.recognize_device(dy, boff)
        frames = n * rec.num_frames(40000)
        got = rec.fetch_labels(n, frames + 48 * n)
        rec.device_free(dx); rec.device_free(dy)
        for w, g in zip(want, got):
            assert np.array_equal(w.view(np.uint8), g.view(np.uint8))
    finally:
        rec.set_wave_format("lin16")
