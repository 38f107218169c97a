"""On-disk forms of the corpus index and the document store."""
import json
import os

import numpy as np
import pytest
import torch

from rag_docvqa_b200 import sharded


def test_corpus_index_files_round_trip_on_cpu(tmp_path):
    g = torch.Generator().manual_seed(1)
    rows = torch.randn(1000, 64, generator=g).to(torch.bfloat16)
    path = str(tmp_path / "idx")
    sharded.save_corpus_index(path, rows, chunk_rows=300)
    idx = sharded.CorpusIndex(path)
    assert (idx.n, idx.d) == (1000, 64) and idx.inv_norm is None
    assert np.array_equal(np.asarray(idx.rows), rows.view(torch.int16).numpy())
    assert json.load(open(os.path.join(path, "meta.json")))["version"] == sharded.INDEX_VERSION
    spans = [sharded.shard_bounds(idx.n, 3, r) for r in range(3)]
    assert sum(b - a for a, b in spans) == 1000
    with open(os.path.join(path, "meta.json"), "w") as f:
        json.dump({"version": 99}, f)
    with pytest.raises(ValueError):
        sharded.CorpusIndex(path)


@pytest.mark.gpu
def test_corpus_index_shards_search_like_the_resident_corpus(tmp_path):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    E = (torch.randn(5000, 128, generator=g) + 0.3).to(dev)
    Q = torch.randn(40, 128, generator=g).to(dev)
    whole = sharded.CorpusShard.from_f32(E)
    path = str(tmp_path / "idx")
    sharded.save_corpus_index(path, whole.rows, whole.inv_norm)
    idx = sharded.CorpusIndex(path)
    ref_v, ref_i = whole.search_local(Q, 10)
    vals, ids = [], []
    for r in range(3):                                   # ranks emulated as slices on one GPU
        shard = idx.load_shard(dev, rank=r, world=3, chunk_rows=700)
        assert torch.equal(shard.rows, whole.rows[shard.id_offset:shard.id_offset + shard.n])
        assert torch.equal(shard.inv_norm, whole.inv_norm[shard.id_offset:shard.id_offset + shard.n])
        v, i = shard.search_local(Q, 10)
        vals.append(v); ids.append(i)
    v, i = sharded.merge_candidates(torch.cat(vals, 1), torch.cat(ids, 1), 10)
    assert torch.equal(i, ref_i) and torch.equal(v, ref_v)


@pytest.mark.gpu
def test_docstore_save_load_gives_the_same_packed_inputs(tmp_path):
    from rag_docvqa_b200 import synth
    from rag_docvqa_b200.docstore import DocStore
    from rag_docvqa_b200.retriever import Retriever
    dev = torch.device("cuda:0")
    batch = synth.make_text_batch("C2", with_lists=True, docs=5, seed=8)
    words = batch["words_text_chunks"]
    table = synth.make_tokens_for_words(words, seed=1)
    store = DocStore.from_lists(words, batch["words_box_chunks"], batch["layout_labels_chunks"], batch["page_indices"],
                                lambda w: table.get(w, [2]), dev, images=batch["images"])
    path = str(tmp_path / "store.npz")
    store.save(path)
    again = DocStore.load(path, dev)
    assert again.B == store.B and sorted(again.host) == sorted(k for k, v in store.host.items() if v is not None)
    for k in again.host:
        assert np.array_equal(again.host[k], store.host[k]), k
    retr = Retriever({"compute_stats": False, "compute_stats_examples": False, "n_stats_examples": 0, "device": "cuda:0", "chunk_num": 5})
    emb, q = [e.to(dev) for e in batch["text_embeddings"]], batch["question_embeddings"].to(dev)
    a, _ = retr.retrieve_packed(emb, q, store, [[5, 6]] * 5)
    b, _ = retr.retrieve_packed(emb, q, again, [[5, 6]] * 5)
    assert torch.equal(a.input_ids, b.input_ids) and torch.equal(a.boxes, b.boxes) and torch.equal(a.hit_rect, b.hit_rect)
