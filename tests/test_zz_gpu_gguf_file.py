"""GGUF file -> device (zgml_b200/host/gguf.py) on the CUDA backend: the loader's two forms give the same decode, and
both match the reference executor.  Composed only of calls the other GPU tests already cover
(`zg_cuda_qweight_upload_gguf`, resident weight descriptors, the decode program); named to sort after them."""
import numpy as np
import pytest

from llama_reference import OracleBackend
from zgml_b200.host import gguf
from zgml_b200.host.gguf import GGUFFile
from zgml_b200.host.llama import DeviceLlamaSession, LlamaConfig

pytestmark = pytest.mark.gpu


def greedy(sess, first, n):
    toks, logs, t = [], [], first
    for _ in range(n):
        lg = sess.step(t).copy()
        t = int(np.argmax(lg))
        toks.append(t)
        logs.append(lg)
    return toks, np.stack(logs)


@pytest.mark.parametrize("kind,tied", [("q8_0", True), ("q4_0", False)])
def test_gguf_file_resident_load_matches_host_load_and_reference(cuda_backend, tmp_path, kind, tied):
    cfg = LlamaConfig(vocab_size=256, d_model=64, n_layers=2, n_heads=4, n_kv_heads=2, d_ff=128, max_seq_len=32, rope_base=5e5,
                      tied_lm_head=tied)
    path = str(tmp_path / "tiny.gguf")
    gguf.write_llama_gguf(path, cfg, kind, seed=5, embed_scale=1.0)
    gf = GGUFFile.open(path)
    assert gguf.config_from_gguf(gf) == cfg
    w_host = gguf.load_direct_quantized(gf)                        # the reference's host form (i8 + f32 scales)
    w_dev, handles = gguf.load_resident(cuda_backend, gf)           # raw block bytes -> HBM, tensor by tensor
    ref = DeviceLlamaSession(OracleBackend(), cfg, w_host)
    a, b = DeviceLlamaSession(cuda_backend, cfg, w_host), DeviceLlamaSession(cuda_backend, cfg, w_dev)
    t_ref, l_ref = greedy(ref, 1, 5)
    t_a, l_a = greedy(a, 1, 5)
    t_b, l_b = greedy(b, 1, 5)
    ref.close(); a.close(); b.close()
    for h in handles:
        h.free()
    assert t_a == t_b and np.array_equal(l_a, l_b)                  # same packed weights either way
    assert t_a == t_ref
    assert np.max(np.abs(l_a - l_ref)) < 1e-3 * np.max(np.abs(l_ref))
