"""S2Chunker with its per-page O(n^2) / O(words x regions) loops on the device.

Drop-in for src._modules.S2Chunker (:1669-1962): same constructor (`config`, `embedder`), same config keys
(`cluster_mode`, `calculate_n_clusters`), same `forward(layout_info, pages_info=None)` -> one cluster array per page,
same helper names.  What the reference computes with Python loops per page runs as two launches for the WHOLE batch:

  * which OCR words lie in which layout region (`containment_ratio > 0.5`, :1720-1731) -- `rdv_layout_assign`, the
    kernel the Chunker uses (float64, bit-exact decisions);
  * the pairwise weight matrices (`_spatial_weights_calculation`, `_semantic_weights_calculation`, `_combined_weights`,
    :1755-1802) -- `rdv_s2_weights` (spatial term float64, bit-exact; cosine term float32 as sklearn computes it).

The spectral clustering that follows (:1815-1857) is the reference's own sklearn calls on the host, unchanged: a few
30 x 30 eigenproblems per page, not a data-parallel path.  Faithful to the reference as written, including: global node
ids that start at len(page words) - 1 in "spatial+semantic" mode with page words (the shadowed loop variable at :1724),
which makes `_add_weights_to_graph` raise IndexError there exactly as the reference does; `max_token_length` is not
set by the constructor (:1678 is commented out), so "heuristic" needs the caller to set it.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .chunker import Chunker
from .functional import _stream_ptr
from .retriever import _device_of


class S2Chunker:
    def __init__(self, config: dict, embedder=None):
        self.config = config
        self.cluster_mode = config.get("cluster_mode", "spatial+semantic")
        self.calculate_n_clusters = config.get("calculate_n_clusters", "heuristic")
        self.graph = None
        self.device = _device_of(config)
        if self.cluster_mode == "spatial+semantic":
            if embedder is None:
                # the reference builds a BiEncoder(config) here (:1682); the encoder is a producer outside this package
                raise ValueError("S2Chunker(cluster_mode='spatial+semantic') needs an embedder with .forward(texts)")
            self.embedder = embedder
            self.tokenizer = getattr(getattr(embedder, "bge_model", None), "tokenizer", None)

    # -- nodes ----------------------------------------------------------------------------------------
    def _nodes_batch(self, layout_info: Sequence[dict], pages_info: Optional[Sequence[Optional[dict]]]):
        """create_nodes_and_edges for a list of pages: ONE rdv_layout_assign launch for every page that needs the
        word -> region membership.  Returns [(nodes, edges, used)] per page."""
        semantic = self.cluster_mode != "spatial"
        jobs, pb, lb, ll = {}, [], [], []
        for p, page in enumerate(layout_info):
            info = pages_info[p] if pages_info is not None else None
            if semantic and info is not None and len(page["boxes"]) and len(info["ocr_tokens"]):
                jobs[p] = len(pb)
                pb.append(np.asarray([b.tolist() if isinstance(b, np.ndarray) else b for b in info["ocr_normalized_boxes"]],
                                     dtype=np.float64).reshape(-1, 4))
                lb.append(np.asarray(page["boxes"], dtype=np.float64).reshape(-1, 4))
                ll.append(np.arange(len(page["boxes"]), dtype=np.int32))
        inside_all = Chunker.assign_words_to_layouts(self, pb, lb, ll)[0] if pb else []
        out = []
        for p, page in enumerate(layout_info):
            boxes, labels = page["boxes"], page["labels"]
            info = pages_info[p] if pages_info is not None else None
            nodes, used, i = [], np.zeros(len(boxes), dtype=bool), 0
            if not semantic or info is None:
                for l, (box, label) in enumerate(zip(boxes, labels)):
                    nodes.append({"global_id": i, "page": 1, "bbox": box, "text": "", "label": label})
                    i += 1
                    used[l] = True
            else:
                words = info["ocr_tokens"]
                if p in jobs:
                    inside = inside_all[jobs[p]]
                    i = len(words) - 1                       # :1724 reuses `i` as the word index
                for l, (box, label) in enumerate(zip(boxes, labels)):
                    members = np.flatnonzero(inside[l]) if p in jobs else ()
                    if len(members) == 0:
                        continue
                    nodes.append({"global_id": i, "page": 1, "bbox": box, "text": " ".join(words[w] for w in members),
                                  "label": label})
                    i += 1
                    used[l] = True
            edges = [(nodes[a]["global_id"], nodes[b]["global_id"]) for a in range(len(nodes)) for b in range(a + 1, len(nodes))]
            out.append((nodes, edges, used))
        return out

    def create_nodes_and_edges(self, page_layout_info: Dict, page_info: Optional[Dict] = None) -> Tuple[List[Dict], List[Tuple], np.ndarray]:
        return self._nodes_batch([page_layout_info], [page_info])[0]

    # -- weights --------------------------------------------------------------------------------------
    def weights_batch(self, boxes_per_page: Sequence[Sequence[Sequence[float]]], embeddings: Optional[Sequence[torch.Tensor]] = None,
                      what: int = _lib.S2_COMBINED) -> List[np.ndarray]:
        """One rdv_s2_weights launch: boxes_per_page[p] (n_p, 4) region boxes, embeddings[p] (n_p, d) fp32 (or None for
        cluster_mode "spatial") -> [n_p x n_p float64 matrices]."""
        P = len(boxes_per_page)
        n = np.asarray([len(b) for b in boxes_per_page], dtype=np.int64)
        node_off = np.zeros(P + 1, dtype=np.int32); np.cumsum(n, out=node_off[1:])
        out_off = np.zeros(P + 1, dtype=np.int64); np.cumsum(n * n, out=out_off[1:])
        N, total = int(node_off[-1]), int(out_off[-1])
        if total == 0:
            return [np.zeros((int(k), int(k))) for k in n]
        dev = self.device
        box = np.concatenate([np.asarray(b, dtype=np.float64).reshape(-1, 4) for b in boxes_per_page])
        parts = [box, out_off, node_off]
        offs, size = [], 0
        for a in parts:
            offs.append(size)
            size += (a.nbytes + 15) // 16 * 16
        with torch.cuda.device(dev):
            host = torch.empty(size, dtype=torch.uint8, pin_memory=True)
            raw = host.numpy()
            for a, o in zip(parts, offs):
                raw[o:o + a.nbytes] = np.frombuffer(a.tobytes(), dtype=np.uint8)
            blob = host.to(dev, non_blocking=True)
            emb, d = None, 0
            if embeddings is not None:
                rows = [e.to(device=dev, dtype=torch.float32) for e in embeddings]
                for p, e in enumerate(rows):
                    if e.dim() != 2 or e.shape[0] != n[p]:
                        # the reference fails here too: (spatial + semantic) / 2 cannot broadcast (:1801)
                        raise ValueError("page %d: %d region boxes but embeddings of shape %s" % (p, n[p], tuple(e.shape)))
                emb = torch.cat(rows).contiguous()
                d = int(emb.shape[1])
            out = torch.empty(total, dtype=torch.float64, device=dev)
            base = blob.data_ptr()
            _lib.check(_lib.lib.rdv_s2_weights(base + offs[0], base + offs[2], P, emb.data_ptr() if emb is not None else None,
                                               d, int(what), base + offs[1], total, out.data_ptr(), _stream_ptr(dev)))
            flat = out.cpu().numpy()
        assert N == len(box)
        return [flat[out_off[p]:out_off[p + 1]].reshape(int(n[p]), int(n[p])).copy() for p in range(P)]

    def _embed(self, nodes: List[Dict]) -> torch.Tensor:
        texts = [node["text"] for node in nodes if node.get("text", "").strip()]        # :1780
        with torch.no_grad():
            return self.embedder.forward(texts)

    def _spatial_weights_calculation(self, nodes: List[Dict]) -> np.ndarray:
        return self.weights_batch([[n["bbox"] for n in nodes]], None, _lib.S2_SPATIAL)[0]

    def _semantic_weights_calculation(self, nodes: List[Dict]) -> np.ndarray:
        emb = self._embed(nodes)
        boxes = [n["bbox"] for n in nodes][:emb.shape[0]]
        return self.weights_batch([boxes], [emb], _lib.S2_SEMANTIC)[0].astype(np.float32)

    def _combined_weights(self, nodes: List[Dict]) -> np.ndarray:
        emb = [self._embed(nodes)] if self.cluster_mode == "spatial+semantic" else None
        return self.weights_batch([[n["bbox"] for n in nodes]], emb)[0]

    # -- clustering (host: the reference's sklearn calls, :1804-1927) ----------------------------------
    def _create_graph(self, nodes: List[int], edges: List[tuple]):
        import networkx as nx
        graph = nx.Graph()
        graph.add_nodes_from(nodes)
        graph.add_edges_from(edges)
        return graph

    def _add_weights_to_graph(self, graph, weights: np.ndarray):
        for u, v in graph.edges():
            graph[u][v]["weight"] = weights[u, v]                                       # indexed by GLOBAL id (:1812)
        return graph

    def _calculate_n_clusters(self, nodes: List[Dict], weights: np.ndarray, min_k: int = 2, max_k: int = 10):
        from sklearn.cluster import KMeans, SpectralClustering
        from sklearn.metrics import silhouette_score
        degree = np.sum(weights, axis=1)
        d_inv_sqrt = np.diag(1.0 / (np.sqrt(degree) + 1e-10))
        l_norm = np.eye(weights.shape[0]) - d_inv_sqrt @ weights @ d_inv_sqrt
        _, eigenvectors = np.linalg.eigh(l_norm)
        embedding = eigenvectors[:, :max_k]
        best_k, best_score, best_labels = min_k, -1, np.full(len(nodes), -1)
        for k in range(min_k, min(max_k, len(nodes) - 1) + 1):
            if self.calculate_n_clusters == "heuristic":
                labels = KMeans(n_clusters=k, random_state=0).fit(embedding).labels_
            elif self.calculate_n_clusters == "best":
                labels = SpectralClustering(n_clusters=k, affinity="precomputed").fit_predict(weights)
            score = silhouette_score(embedding, labels)
            if score > best_score:
                best_score, best_k, best_labels = score, k, labels
        return best_k, best_labels

    def _cluster_graph(self, graph, weights: np.ndarray, n_clusters: int = 3) -> Dict[int, int]:
        from sklearn.cluster import SpectralClustering
        labels = SpectralClustering(n_clusters=n_clusters, affinity="precomputed").fit_predict(weights)
        return {node: label for node, label in zip(graph.nodes(), labels)}

    def _group_nodes_by_cluster(self, clusters: Dict[int, int]) -> Dict[int, List[int]]:
        groups: Dict[int, List[int]] = {}
        for node, cluster_id in clusters.items():
            groups.setdefault(cluster_id, []).append(node)
        return groups

    def _split_clusters_by_token_length(self, clusters: Dict[int, int], nodes: List[Dict]) -> Dict[int, int]:
        by_id = {}
        for n in nodes:
            by_id.setdefault(n["global_id"], n)
        updated, counter = {}, 0
        for _cluster_id, node_ids in self._group_nodes_by_cluster(clusters).items():
            current, length = [], 0
            for node_id in node_ids:
                node = by_id.get(node_id)
                if not node:
                    continue
                n_tokens = len(self.tokenizer.tokenize(node["text"]))
                if length + n_tokens > self.max_token_length:
                    for member in current:
                        updated[member] = counter
                    counter += 1
                    current, length = [], 0
                current.append(node_id)
                length += n_tokens
            for member in current:
                updated[member] = counter
            counter += 1
        return updated

    def _cluster_with_weights(self, nodes: List[Dict], edges: List[tuple], weights: np.ndarray) -> Dict[int, int]:
        graph = self._create_graph([node["global_id"] for node in nodes], edges)
        self.graph = self._add_weights_to_graph(graph, weights)
        n_clusters, best_labels = self._calculate_n_clusters(nodes, weights)
        if self.calculate_n_clusters == "heuristic":
            clusters = self._cluster_graph(self.graph, weights, n_clusters)
            return self._split_clusters_by_token_length(clusters, nodes)
        return {node: label for node, label in zip(self.graph.nodes(), best_labels)}

    def cluster(self, nodes: List[Dict], edges: List[tuple]) -> Dict[int, int]:
        return self._cluster_with_weights(nodes, edges, self._combined_weights(nodes))

    def forward(self, layout_info: List[Dict], pages_info: Optional[List[Dict]] = None) -> List[np.ndarray]:
        built = self._nodes_batch(layout_info, pages_info)
        todo = [p for p, page in enumerate(layout_info) if len(page["boxes"]) and len(built[p][0]) >= 2]
        emb = [self._embed(built[p][0]) for p in todo] if self.cluster_mode == "spatial+semantic" else None
        weights = self.weights_batch([[n["bbox"] for n in built[p][0]] for p in todo], emb)      # ONE launch for the batch
        weights = dict(zip(todo, weights))
        batch_clusters = []
        for p, page in enumerate(layout_info):
            if len(page["boxes"]) == 0:
                batch_clusters.append(np.array([]))
                continue
            nodes, edges, used = built[p]
            if len(nodes) < 2:
                batch_clusters.append(np.full(len(page["boxes"]), -1))
                continue
            clusters = self._cluster_with_weights(nodes, edges, weights[p])
            clusters = [label for _, label in sorted(clusters.items(), key=lambda item: item[0])]
            complete = np.full(len(used), -1)
            complete[used] = clusters
            batch_clusters.append(complete)
        return batch_clusters
