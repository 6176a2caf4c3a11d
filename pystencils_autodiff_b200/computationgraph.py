"""Dependency graph of a recorded call queue.

The reference's ``ComputationGraph`` (/root/reference/src/pystencils_autodiff/computationgraph.py:17-163) turns the
``call_queue`` of a ``GraphDataHandling`` into array-version and computation nodes and draws them with graphviz.  This one
does the same for the queue ``SlabDataHandling`` records — ``KernelCall``, ``Communication`` (the ghost-plane exchange),
``Swap`` — numbering every write of an array (``u #0 -> u #1 ...``), and additionally answers the scheduling question
the picture is for: which recorded calls do not depend on each other (``levels()``).  ``to_dot()`` returns DOT text; no
graphviz package is needed.
"""
from collections import OrderedDict

__all__ = ['ComputationGraph']


class ComputationGraph:
    class Node:
        def __init__(self, index, kind, label):
            self.index, self.kind, self.label = index, kind, label
            self.inputs, self.outputs = [], []          # array snapshots ("u #1")

        def __repr__(self):
            return '%s[%d] %s' % (self.kind, self.index, self.label)

    def __init__(self, data_handling_or_queue, kernel_io=None):
        """``data_handling_or_queue``: a ``SlabDataHandling`` (its ``call_queue`` and the read / written fields of the
        kernels it has run) or a queue plus ``kernel_io = {kernel name: (read field names, written field names)}``."""
        if hasattr(data_handling_or_queue, 'call_queue'):
            kernel_io = dict(getattr(data_handling_or_queue, 'kernel_io', {}), **(kernel_io or {}))
            queue = data_handling_or_queue.call_queue
        else:
            queue = data_handling_or_queue
        # a ``TimeloopRun`` contributes the calls of ONE of its steps (the reference nests a sub-graph of the loop's
        # ``_single_step_asts``, computationgraph.py:71-74, sharing the write counter: the same versions result)
        self.call_list = []
        for c in queue:
            if isinstance(c, tuple) and c and c[0] == 'TimeloopRun':
                self.call_list.extend(c[2])
            else:
                self.call_list.append(c)
        self.kernel_io = dict(kernel_io or {})
        self.write_counter = {}
        self.reads = OrderedDict()          # snapshot -> nodes reading it
        self.writes = OrderedDict()         # snapshot -> node that produced it
        self.computation_nodes = []
        for i, c in enumerate(self.call_list):
            kind = c[0] if isinstance(c, tuple) else str(c)
            if kind in ('KernelCall', 'KernelCall+Swap'):
                name = c[1]
                if name not in self.kernel_io:
                    raise KeyError('no read / write information for kernel %r: pass kernel_io' % name)
                r, w = self.kernel_io[name]
                node = self._add(i, 'kernel', name, r, w)
                if kind == 'KernelCall+Swap':
                    for a, b in c[-1]:
                        self._add(i, 'swap', 'Swap %s <-> %s' % (a, b), [a, b], [a, b])
                del node
            elif kind == 'Swap':
                self._add(i, 'swap', 'Swap %s <-> %s' % (c[1], c[2]), [c[1], c[2]], [c[1], c[2]])
            elif kind == 'Communication':
                self._add(i, 'communication', 'ghost planes of %s' % c[1], [c[1]], [c[1]])
            elif kind == 'Fill':
                self._add(i, 'fill', 'Fill %s' % c[1], [], [c[1]])
            elif kind == 'DataTransfer':
                self._add(i, 'transfer', '%s %s' % (c[2], c[1]), [c[1]], [c[1]] if c[2] == 'HOST_TO_DEVICE' else [])
            # markers without data dependencies (FieldOutput, GhostTensorExtraction) read only
            elif kind in ('FieldOutput',):
                self._add(i, 'output', 'save %s' % ', '.join(c[1]), list(c[1]), [])
            elif kind == 'GhostTensorExtraction':
                self._add(i, 'output', 'extract %s' % c[1], [c[1]], [])

    def _snapshot(self, name):
        return '%s #%d' % (name, self.write_counter.get(name, 0))

    def _add(self, index, kind, label, reads, writes):
        node = self.Node(index, kind, label)
        for name in reads:
            snap = self._snapshot(name)
            node.inputs.append(snap)
            self.reads.setdefault(snap, []).append(node)
        for name in writes:
            self.write_counter[name] = self.write_counter.get(name, 0) + 1
            snap = self._snapshot(name)
            node.outputs.append(snap)
            self.writes[snap] = node
        self.computation_nodes.append(node)
        return node

    # -- scheduling view ------------------------------------------------------------------------------------------
    def dependencies(self):
        """node -> set of nodes it must run after (true, anti and output dependencies on array versions)."""
        deps = {n: set() for n in self.computation_nodes}
        last_write, readers = {}, {}
        for n in self.computation_nodes:
            for snap in n.inputs:
                name = snap.rsplit(' #', 1)[0]
                if name in last_write:
                    deps[n].add(last_write[name])                 # read after write
                readers.setdefault(name, []).append(n)
            for snap in n.outputs:
                name = snap.rsplit(' #', 1)[0]
                if name in last_write:
                    deps[n].add(last_write[name])                 # write after write
                deps[n].update(r for r in readers.get(name, []) if r is not n)     # write after read
                last_write[name] = n
                readers[name] = []
        for n in deps:
            deps[n].discard(n)
        return deps

    def levels(self):
        """The recorded calls grouped into levels: every call depends only on calls of earlier levels, so the calls of one
        level may run concurrently (e.g. on different streams)."""
        deps = self.dependencies()
        level = {}
        for n in self.computation_nodes:
            level[n] = 1 + max((level[d] for d in deps[n]), default=-1)
        out = [[] for _ in range(1 + max(level.values(), default=-1))]
        for n in self.computation_nodes:
            out[level[n]].append(n)
        return out

    # -- drawing --------------------------------------------------------------------------------------------------
    _COLOURS = {'kernel': '#0056db', 'swap': '#ff5600', 'communication': '#00a070', 'fill': '#888888',
                'transfer': '#888888', 'output': '#888888'}

    def to_dot(self, graph_style=None, with_code=False):
        lines = ['digraph "%d" {' % id(self)]
        for k, v in (graph_style or {}).items():
            lines.append('  %s=%s;' % (k, v))
        snaps = list(OrderedDict.fromkeys(list(self.reads) + list(self.writes)))
        for s_ in snaps:
            lines.append('  "%s" [style=filled, fillcolor="#a056db"];' % s_)
        for n in self.computation_nodes:
            nid = 'n%d_%d' % (n.index, self.computation_nodes.index(n))
            lines.append('  %s [style=filled, fillcolor="%s", label="%s"];' % (nid, self._COLOURS[n.kind], n.label))
            for s_ in n.inputs:
                lines.append('  "%s" -> %s;' % (s_, nid))
            for s_ in n.outputs:
                lines.append('  %s -> "%s";' % (nid, s_))
        lines.append('}')
        return '\n'.join(lines)

    def to_dot_file(self, path, graph_style=None, with_code=False):
        with open(path, 'w') as fh:
            fh.write(self.to_dot(graph_style, with_code))
