"""TMA-ring score kernel at C2 / C3 for different stage heights (tile_rows)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_docvqa_b200 import functional as F, synth, _lib
dev = torch.device("cuda:0")
for wl, R in (("C2", 10), ("C3", 2)):
    w = synth.WORKLOADS[wl]
    batches = [synth.make_text_batch(wl, device=dev, emb_seed=1000 * (r + 1)) for r in range(R)]
    nbytes = sum(e.numel() * 4 for e in batches[0]["text_embeddings"])
    for algo, rows_list in ((1, (0,)), (2, (2, 4, 8, 16))):
        for tr in rows_list:
            try:
                tabs = [F.build_doc_table(b["text_embeddings"], w.dim, dev, tile_rows=tr, algo=algo) for b in batches]
            except Exception as e:
                print(wl, algo, tr, "rejected:", str(e)[:80]); continue
            qs = [b["question_embeddings"] for b in batches]
            try:
                for i in range(5): F.score_table(tabs[i % R], qs[i % R])
            except Exception as e:
                print(wl, algo, tr, "rejected:", str(e)[:80]); continue
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 100 if wl == "C2" else 10
            e0.record()
            for i in range(n): F.score_table(tabs[i % R], qs[i % R])
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / n * 1e3
            print("%s algo=%d tile_rows=%d (tiles %d): %.2f us  %.0f GB/s" % (wl, algo, tabs[0].tile_rows, tabs[0].total_tiles, us, nbytes / us / 1e3))
    del batches
    torch.cuda.empty_cache()
