"""Row-partitioned propagation over several GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Output rows of a layer
are independent given all input rows, so rank r owns a block of user rows and a block of item
(+property) rows - interleaving the two node types balances edges, because user rows are
short and item rows long - computes them with the same kernels on CSR row slices (global
column ids), and the layer outputs are exchanged with an all-gather.  Every row's reduction
order is rank-independent (ascending columns, fixed chunk tree), so the G-rank result equals
the 1-rank result bit for bit.  Scoring is sharded by user with items replicated: no
collective.
"""
import torch
import torch.distributed as dist


def block_ranges(n_rows_by_type, world_size):
    """[[(r0, r1) per node type] per rank]; each type's rows are cut into world_size blocks of
    ceil(n/world_size) rows (the last may be short or empty).  `n_rows_by_type` lists the
    sizes of the consecutive id ranges (users, items[, properties])."""
    out = [[] for _ in range(world_size)]
    base = 0
    for n in n_rows_by_type:
        blk = -(-n // world_size)
        for r in range(world_size):
            lo = min(r * blk, n)
            hi = min(lo + blk, n)
            out[r].append((base + lo, base + hi))
        base += n
    return out


def exchange_rows(x, ranges, group=None):
    """All ranks end up with every row block of x (each block is authored by its owner).

    Even blocks take one all_gather_into_tensor per node type (in place: the owner's block is
    already where the gathered tensor wants it); ragged blocks fall back to one broadcast per
    owner.  x must be row-contiguous [N, H]."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_types = len(ranges[0])
    for t in range(n_types):
        sizes = [ranges[r][t][1] - ranges[r][t][0] for r in range(world)]
        lo = ranges[0][t][0]
        if len(set(sizes)) == 1 and sizes[0] > 0 and x.is_contiguous():
            span = x[lo:lo + world * sizes[0]]
            mine = x[ranges[rank][t][0]:ranges[rank][t][1]]
            dist.all_gather_into_tensor(span, mine, group=group)
        else:
            for r in range(world):
                a, b = ranges[r][t]
                if b > a:
                    dist.broadcast(x[a:b], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
    return x


class RowPartition:
    """Attach to a SequentialGNN to run its layer loop on this rank's row blocks.

    What travels per layer is the [N, H] operand the sparse kernel gathers from:
      GCN / RGCN : Z = X W is computed for OWN rows only, all-gathered, then A_hat Z on own rows
                   (the dense transform is never replicated; X itself is never exchanged);
      GAT        : z = X W, q = z.a_neigh for own rows, all-gathered; p stays local;
      GraphSage / LightGCN : the layer INPUT X^(l) is all-gathered (layer 0 reads the
                   replicated embedding table).
    After the last layer only the node types in `final_types` (default: everything but users)
    are exchanged, because scoring is sharded by user."""

    def __init__(self, n_rows_by_type, group=None, final_types=None, col_splits=4):
        self.group = group
        import os
        self.col_splits = int(os.environ.get("CBRS_COL_SPLITS", col_splits))  # GCN: the all-gather of column block s+1 overlaps the SpMM of block s
        self._comm = None
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.ranges = block_ranges(n_rows_by_type, self.world)
        self.mine = [rg for rg in self.ranges[self.rank] if rg[1] > rg[0]]
        self.final_types = list(range(1, len(n_rows_by_type))) if final_types is None else list(final_types)
        self._slices = {}
        self._bufs = {}

    def attach(self, seq_gnn):
        seq_gnn.partition = self
        return self

    def csr_slices(self, view_name, graph):
        if view_name not in self._slices:
            full = getattr(graph, view_name)
            self._slices[view_name] = [full.row_slice(a, b) for a, b in self.mine]
        return self._slices[view_name]

    def release_full_views(self, graph):
        """Keep only this rank's row slices in HBM."""
        graph._views.clear()
        graph.release_coo()

    def local_edges(self, view_name):
        return sum(s.nnz for s in self._slices.get(view_name, []))

    def _buf(self, key, n, w, device):
        if key not in self._bufs:
            self._bufs[key] = torch.empty(n, w, dtype=torch.float32, device=device)
        return self._bufs[key]

    def _exchange(self, x, types=None):
        from . import ops
        ranges = self.ranges if types is None else [[rg[t] for t in types] for rg in self.ranges]
        if ops.PROFILE_ON:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            exchange_rows(x, ranges, self.group)
            e1.record()
            ops.PROFILE.append(("exchange", e0, e1, x.numel() * 4))
            return x
        return exchange_rows(x, ranges, self.group)

    def _pipelined_ok(self, layer):
        return self.col_splits > 1 and layer.channels % (4 * self.col_splits) == 0

    def _gcn_pipelined(self, layer, l, x_full, out, graph, relu):
        """GCN layer with the exchange hidden behind the sparse kernel: Z = X W is produced and
        all-gathered in `col_splits` column blocks on a side stream, and the SpMM of block s runs
        while block s+1 is still on the wire.  Columns are independent, so the bits are the same
        as the unsplit layer's."""
        from . import ops
        n, dev = x_full.shape[0], x_full.device
        S, hs = self.col_splits, layer.channels // self.col_splits
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        zs = [self._buf(("zs", l, s), n, hs, dev) for s in range(S)]
        ready = []
        for s in range(S):
            w_s = layer.kernel[:, s * hs:(s + 1) * hs].contiguous()
            for a, b in self.mine:
                ops.dense(x_full[a:b], w_s, out=zs[s][a:b])
            produced = torch.cuda.Event()
            produced.record(main)
            with torch.cuda.stream(self._comm):
                self._comm.wait_event(produced)
                self._exchange(zs[s])
                done = torch.cuda.Event()
                done.record(self._comm)
            ready.append(done)
        for s in range(S):
            main.wait_event(ready[s])
            bias = layer.bias[s * hs:(s + 1) * hs] if layer.bias is not None else None
            for sl in self.csr_slices("norm", graph):
                ops.spmm(sl, zs[s], out[sl.row_offset:sl.row_offset + sl.n_rows, s * hs:(s + 1) * hs], bias=bias, relu=relu)

    def propagate(self, seq):
        """SequentialGNN.call on the partition.  Returns [N, D_out]; rows of other ranks' users
        are NOT valid unless final_types covers type 0."""
        from . import _lib as L
        from . import ops
        from .layers import GATConv, GCNConv, GraphSageConv, LightGCNConv, RGCNConv
        emb = seq.embeddings
        n, dev = emb.shape[0], emb.device
        widths = seq._widths()
        graph = seq.adj_matrix
        x_full, x_full_valid = emb, True   # the layer input; valid on all rows?
        hs = [emb]
        # with 'concatenation' and layers that exchange their own operand (GCN/RGCN/GAT) the outputs go
        # straight into column slices of one [N, D_out] buffer, as on a single GPU
        concat = seq.final_node == 'concatenation' and all(isinstance(ly, (GCNConv, RGCNConv, GATConv))
                                                           for ly in seq.seq_layers)
        if concat:
            cbuf = self._buf(("concat",), n, sum(widths), dev)
            for a, b in self.mine:
                cbuf[a:b, :widths[0]].copy_(emb[a:b])
            off = widths[0]
        for l, layer in enumerate(seq.seq_layers):
            if not layer.built:
                layer.build([(n, widths[l]), None])
                layer.built = True
            if concat:
                out = cbuf[:, off:off + widths[l + 1]]
                off += widths[l + 1]
            else:
                out = self._buf(("h", l), n, widths[l + 1], dev)  # own rows valid
            relu = getattr(layer, "activation", None) == "relu"
            if isinstance(layer, GCNConv) and self._pipelined_ok(layer):
                self._gcn_pipelined(layer, l, x_full, out, graph, relu)
            elif isinstance(layer, (GCNConv, RGCNConv)):
                kernels = layer.kernels if isinstance(layer, RGCNConv) else [layer.kernel]
                z = self._buf(("z", l), len(kernels) * n, layer.channels, dev)
                for r, w in enumerate(kernels):
                    for a, b in self.mine:
                        ops.dense(x_full[a:b], w, out=z[r * n + a:r * n + b])
                    self._exchange(z[r * n:(r + 1) * n])
                for sl in self.csr_slices("norm", graph):
                    ops.spmm(sl, z, out[sl.row_offset:sl.row_offset + sl.n_rows], bias=layer.bias, relu=relu)
            elif isinstance(layer, GATConv):
                z = self._buf(("z", l), n, layer.channels, dev)
                p = self._buf(("p", l), n, 1, dev)
                q = self._buf(("q", l), n, 1, dev)
                for a, b in self.mine:
                    _, pp, qq = ops.dense(x_full[a:b], layer.kernel.reshape(widths[l], layer.channels),
                                          rowop=L.ROWOP_ATTN, a_self=layer.attn_kernel_self.reshape(-1),
                                          a_neigh=layer.attn_kernel_neighs.reshape(-1), out=z[a:b])
                    p[a:b, 0].copy_(pp)
                    q[a:b, 0].copy_(qq)
                self._exchange(z)
                self._exchange(q)
                for sl in self.csr_slices("raw", graph):
                    ops.gat(sl, z, p.reshape(-1), q.reshape(-1), out[sl.row_offset:sl.row_offset + sl.n_rows],
                            bias=layer.bias, relu=relu, row_offset=sl.row_offset)
            elif isinstance(layer, (GraphSageConv, LightGCNConv)):
                if not x_full_valid:
                    self._exchange(x_full)
                view = "raw" if isinstance(layer, GraphSageConv) else "norm"
                for sl in self.csr_slices(view, graph):
                    layer([x_full, graph], out=out[sl.row_offset:sl.row_offset + sl.n_rows], csr=sl)
            else:
                raise NotImplementedError("no partitioned form for {}".format(type(layer).__name__))
            x_full, x_full_valid = out, False
            hs.append(out)
        red = cbuf if concat else seq.reduce(hs)
        if self.final_types:
            red = red if red.is_contiguous() else red.contiguous()
            self._exchange(red, self.final_types)
        return red
