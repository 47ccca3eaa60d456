"""Static non-zero lists of the partitioned adjacency (the "SpA" arguments of the C ABI).

A_eff = A*imp (+ A2*imp2 + A3*imp3) keeps the fixed zero pattern of A U A2 U A3 because the
importances multiply element-wise (net/st_gcnold.py:86, net/st_gcn_msgcn.py:116-117;
SURVEY.md App. A: 197 of 2500 entries for NTU 'spatial_3_sym').  The lists are built once per
model from the registered buffers and cached on the device."""
import numpy as np
import torch


class SparsePattern(object):
    """Index arrays for a (K, V, V) stack with non-zero pattern ``mask``.

    canonical order: sorted by (k, v, w) -> ``flat_idx`` (index into A_eff.reshape(-1))
    destination order (forward aggregation): grouped by (k, w): dst_ptr[K*V+1], dst_src, dst_id
    source order (backward): grouped by v: src_ptr[V+1], src_kw (= k*V + w), src_id
    """

    def __init__(self, mask, device):
        mask = np.asarray(mask, dtype=bool)
        K, V, V2 = mask.shape
        assert V == V2
        k, v, w = np.nonzero(mask)                      # already sorted by (k, v, w)
        nnz = k.size
        self.K, self.V, self.nnz = K, V, int(nnz)
        ids = np.arange(nnz)
        order_d = np.lexsort((v, w, k))                 # by k, then w, then v
        counts = np.bincount(k[order_d] * V + w[order_d], minlength=K * V)
        dst_ptr = np.concatenate([[0], np.cumsum(counts)])
        order_s = np.lexsort((w, k, v))                 # by v, then k, then w
        counts_s = np.bincount(v[order_s], minlength=V)
        src_ptr = np.concatenate([[0], np.cumsum(counts_s)])

        def dev(a, dt=torch.int32):
            return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(device)
        self.flat_idx = dev(k * V * V + v * V + w, torch.int64)
        inv = np.full(K * V * V, -1, dtype=np.int64)
        inv[k * V * V + v * V + w] = ids
        self.inv_idx = dev(inv)                         # [K*V*V] -> canonical id of the entry, or -1
        self.id_kw = dev(k * V + w)                     # [nnz] -> k*V + w (destination column) of entry id
        self.dst_ptr, self.dst_src, self.dst_id = dev(dst_ptr), dev(v[order_d]), dev(ids[order_d])
        # transposed adjacency for the input gradient: grouped by (k, v), "source" = w
        order_t = np.lexsort((w, v, k))                 # by k, then v, then w
        counts_t = np.bincount(k[order_t] * V + v[order_t], minlength=K * V)
        self.t_ptr = dev(np.concatenate([[0], np.cumsum(counts_t)]))
        self.t_src, self.t_id = dev(w[order_t]), dev(ids[order_t])
        self.src_ptr = dev(src_ptr)
        self.src_kw = dev(k[order_s] * V + w[order_s])
        self.src_id = dev(ids[order_s])
        # joint pairs (v, w) that are non-zero in SOME partition, grouped by destination w
        # (csrc/gcn_pair_tc.cu): pair id = position in the (w, v)-sorted list
        key = w.astype(np.int64) * V + v
        uniq, inverse = np.unique(key, return_inverse=True)
        self.npairs = int(uniq.size)
        self._pair_w = (uniq // V).astype(np.int64)
        self._pair_v_host = (uniq % V).astype(np.int64)
        self.pair_v = dev(uniq % V)                      # [npairs] source joint of every pair
        pair_of = np.full(V * V, -1, dtype=np.int64)
        pair_of[self._pair_v_host * V + self._pair_w] = np.arange(uniq.size)
        self.pair_of = dev(pair_of)                      # [V*V] (v*V + w) -> pair id, -1 outside the pattern
        self.entry_pair = dev(inverse)                   # [nnz] canonical entry -> pair id
        self.k_ptr = dev(np.concatenate([[0], np.cumsum(np.bincount(k, minlength=K))]))
        self._device = device
        self._pair_items = {}

    def _cover(self, nd_max, ns_max):
        """Greedy cover of the joint-pair pattern by (sources x destinations) blocks of at most
        ns_max x nd_max joints: [(sources, destinations, owned-cell mask)], every pattern pair owned by
        exactly one block.  The skeleton patterns are block-structured (limbs), so the blocks fill up."""
        V = self.V
        rem = np.zeros((V, V), dtype=bool)                 # [v][w]
        rem[np.asarray(self._pair_v_host), np.asarray(self._pair_w)] = True
        blocks = []
        while rem.any():
            D = [int(rem.sum(0).argmax())]
            while len(D) < nd_max:                         # destinations sharing the most sources with D
                share = [(int((rem[:, w] & rem[:, D].any(1)).sum()), -w) for w in range(V) if w not in D]
                c, w = max(share)
                if c == 0:
                    break
                D.append(-w)
            cnt = rem[:, D].sum(1)
            S = [int(v) for v in np.argsort(-cnt, kind='stable')[:ns_max] if cnt[v] > 0]
            mask = 0
            for vi, v in enumerate(S):
                for jd, w in enumerate(D):
                    if rem[v, w]:
                        mask |= 1 << (vi * len(D) + jd)
                        rem[v, w] = False
            blocks.append((S, D, mask))
        return blocks

    def pair_items(self, cin, cout):
        """Work items of istgcn_gcn_pair_grads for this channel pair -> (items int32 [n][8], ctas int32
        [m][4], joints int32 [...]): blocks of the pair pattern x column chunks, and the thread blocks
        dealt to them (include/istgcn_b200.h)."""
        hit = self._pair_items.get((cin, cout))
        if hit is None:
            import os
            # several destinations per item pay off when an M-block is one whole source joint
            # (cin >= 128: 128 -> 128 at 2 x 2 joints: 371 -> 317 us); with half-block sources the cover
            # wastes too many cells of its blocks (64 -> 64 at 4 x 4: 0.63x the operand bytes, 1.9x the
            # MMAs, measured 284 -> 296 us).  ISTGCN_PAIR_BLOCKS=0 / 1 forces one / several destinations.
            env = os.environ.get('ISTGCN_PAIR_BLOCKS')
            single = env == '0' or (env != '1' and cin < 128)
            # candidates: ncw output channels of nd destinations side by side (N = nd * ncw <= 256), ns
            # sources stacked (M-blocks * N <= 512 columns of tensor memory); cost = operand rows loaded
            # per K-tile = ns * cin + nd * ncw per block and column chunk (the kernel is L2-stream bound)
            best = None
            for ncw in sorted({c for c in (256, 128, 96, 64, 32) if cout % c == 0}, reverse=True):
                for nd in (1, 2, 3, 4):
                    n = nd * ncw
                    if n > 256 or (single and nd > 1):
                        continue
                    ns = min(16, (512 // n) * 128 // cin, 32 // nd)
                    if ns < 1:
                        continue
                    blocks = self._cover(nd, ns)
                    cost = sum(len(S) * cin + len(D) * ncw for S, D, _ in blocks) * (cout // ncw)
                    if best is None or cost < best[0]:
                        best = (cost, ncw, blocks)
            _, ncw, blocks = best
            joints, rows, costs = [], [], []
            for S, D, mask in blocks:
                d0 = len(joints)
                joints += D
                s0 = len(joints)
                joints += S
                for col0 in range(0, cout, ncw):
                    rows.append((d0, len(D), s0, len(S), col0, ncw, mask, 0))
                    costs.append(len(S) * cin + len(D) * ncw)
            items = torch.as_tensor(np.asarray(rows, dtype=np.int64).astype(np.int32).reshape(-1, 8)).to(self._device)
            joints_t = torch.as_tensor(np.asarray(joints, dtype=np.int32)).to(self._device)
            # thread blocks per item in proportion to its cost, about one block per SM in total; block j
            # of an item with s blocks takes the K-tiles j, j + s, j + 2s, ...
            sms = torch.cuda.get_device_properties(self._device).multi_processor_count \
                if torch.cuda.is_available() else 148
            cost = np.asarray(costs, dtype=np.float64)
            # greedy apportionment (the next block goes to the slowest item) over 1 .. 4 waves of
            # thread blocks: the fewest waves whose slowest block is within 20 % of the mean
            for waves in range(max(1, -(-len(rows) // sms)), 5):
                share = np.ones(len(rows), dtype=int)
                for _ in range(max(0, sms * waves - len(rows))):
                    share[int(np.argmax(cost / share))] += 1
                if (cost / share).max() <= 1.2 * cost.sum() / share.sum():
                    break
            ctas = [(i, j, int(s_), 0) for i, s_ in enumerate(share) for j in range(int(s_))]
            ctas = torch.as_tensor(np.asarray(ctas, dtype=np.int32).reshape(-1, 4)).to(self._device)
            hit = self._pair_items[(cin, cout)] = (items, ctas, joints_t)
        return hit

    @classmethod
    def identity(cls, V, device):
        """K = 1, A = I: turns the fused graph conv into a plain (strided) 1x1 convolution."""
        return cls(np.eye(V, dtype=bool)[None], device)
