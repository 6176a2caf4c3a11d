"""``CompiledKernel`` subclasses whose launch is the CPU replay of the emitted kernel (tests/march_emulator.py) on CPU
tensors.  TEST INFRASTRUCTURE: lets the host logic above the launch (data handling, time loops, autograd Functions) run
without a GPU while the per-cell arithmetic still comes from the product's emitted CUDA source."""
import march_emulator as emu
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel


class ReplayKernel(CompiledKernel):
    """Instance selection mirrors ``CompiledKernel.__call__``: march kernels for dense 16-byte aligned rows (mask-free
    instance when every written cell is evaluated), else the generic kernel; ``_variant='march_x2'`` = fused pair."""
    launches = []

    def __call__(self, *, _range=None, _variant=None, _stream=None, **kwargs):
        import fake_cuda
        if fake_cuda.FakeGraph.capturing is not None:       # a capturing stream records the launch and executes nothing
            graph, fake_cuda.FakeGraph.capturing = fake_cuda.FakeGraph.capturing, None
            graph.ops.append(lambda kw=dict(kwargs): self(_range=_range, _variant=_variant, **kw))
            fake_cuda.FakeGraph.capturing = graph
            return None
        nd = self.ir.ndim
        tensors = [kwargs[f.name] for f in self.fields]
        scal = [float(kwargs[s_]) for s_ in self.scalars]
        variant = _variant or self._select_variant(tensors)
        if variant.startswith('march') and self._components:
            import pytest
            pytest.skip('replay: structure-of-arrays component kernels are not modelled (per-component pointers)')
        if variant == 'march_x2':
            ek = self.emitted('march_x2')
        elif variant == 'march':
            if _range is not None:
                same = (list(_range['iter_lo'][:nd]) == list(_range['write_lo'][:nd]) and
                        list(_range['iter_hi'][:nd]) == list(_range['write_hi'][:nd]))
            else:
                same = self.ir.boundary == 'zeros' or self.ir.ghost_layers == 0
            ek = self._emitted['march_nomask' if same else 'march']
        else:
            ek = self._emitted['generic']
        type(self).launches.append(ek.name)
        rng = None if _range is None else {k: v for k, v in _range.items() if not k.startswith('_')}
        arrays = [t.detach().numpy() for t in tensors]
        if variant == 'march_x2' and nd == 2:      # the lifted pair takes one-plane 3-D fields, like CompiledKernel.__call__
            arrays = [a[None] for a in arrays]
        if variant == 'generic':
            emu.run_generic(ek, arrays, scal, launch_range=rng)
        else:
            emu.run(ek, arrays, scal, launch_range=rng)
        self.last_variant = 'march' if variant.startswith('march') else variant
        self.last_instance = variant if variant != 'march' else \
            ('march_nomask' if ek is self._emitted.get('march_nomask') else 'march')
