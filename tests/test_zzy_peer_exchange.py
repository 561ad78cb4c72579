"""GPU: the peer-memory variants of the sharded exchanges (kb_*_peers producers store into every shard's buffers,
PeerShardedActiveWindow + LocalPeers: shards as handles on one device). Written after round 1's GPU minutes were spent
— green under tools/cuda_emu, never run on hardware — so it sorts after the hardware-validated files (the round-end
run uses -x)."""
import pytest

from test_sharded_pipeline import run_sharded_vs_unsharded

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nshards", [2, 4])
def test_product_shards_peer_memory_exchange(oracle_lib, product_lib, nshards):
    run_sharded_vs_unsharded(oracle_lib, product_lib, "kb_", nshards, "cuda", sep=2.0, peers=True)
