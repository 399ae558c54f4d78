"""ExponentialMovingAverage with the reference's surface (lib/algorithms/ema.py:10-98).  The shadow parameters live in one
flat buffer; when the live parameters are flat too (losses.FlatAdam) ``update`` is ONE native kernel
(``dpb_ema_update``), otherwise one per parameter.  ``state_dict`` keeps the reference's layout
(decay / num_updates / shadow_params as a list of per-parameter tensors), so checkpoints are interchangeable."""
import torch

from . import _lib as L


class ExponentialMovingAverage:
    def __init__(self, parameters, decay=0.999, use_num_updates=True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError('Decay must be between 0 and 1')
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        params = [p for p in parameters if p.requires_grad]
        self._alloc(params)
        self.collected_params = []

    def _alloc(self, params, values=None):
        n = sum(p.numel() for p in params)
        dev = params[0].device if params else 'cpu'
        self._flat = torch.empty(n, device=dev)
        self.shadow_params, off = [], 0
        src = values if values is not None else params
        with torch.no_grad():
            for p, s in zip(params, src):
                v = self._flat[off:off + p.numel()].view_as(p)
                v.copy_(s.detach())
                self.shadow_params.append(v)
                off += p.numel()

    @staticmethod
    def _flat_view(params):
        """The flat buffer behind ``params`` if they are consecutive views of one (losses.FlatAdam), else None."""
        if not params:
            return None
        base = params[0]
        ptr = base.data_ptr()
        for p in params:
            if p.data_ptr() != ptr or not p.is_contiguous():
                return None
            ptr += p.numel() * 4
        n = sum(p.numel() for p in params)
        st = base.untyped_storage()
        if (base.data_ptr() - st.data_ptr()) + n * 4 > st.nbytes():
            return None
        return torch.as_strided(base.detach(), (n,), (1,))

    def next_one_minus_decay(self):
        """1 - decay of the NEXT update (ema.py:44-47), without advancing the counter."""
        decay = self.decay
        if self.num_updates is not None:
            decay = min(decay, (2 + self.num_updates) / (11 + self.num_updates))
        return 1.0 - decay

    def update(self, parameters, omd_dev=None):
        """ema.py:35-50: s -= (1 - decay) (s - p), decay capped by (1 + n) / (10 + n).  ``omd_dev`` (device fp32 [1]): the
        kernel reads 1 - decay from device memory (CUDA-graph replay; the caller filled it with next_one_minus_decay())."""
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        omd = 1.0 - decay
        params = [p for p in parameters if p.requires_grad]
        if not params:
            return
        L.require_cuda(params[0], 'parameters')
        lib = L.load()
        flat = self._flat_view(params) if params[0].dtype == torch.float32 else None
        with torch.no_grad():
            if flat is not None and flat.numel() == self._flat.numel():
                L.check(lib.dpb_ema_update(L.ptr(self._flat), L.ptr(flat), flat.numel(), float(omd), L.ptr(omd_dev),
                                           L.current_stream(flat.device)))
            else:
                for s, p in zip(self.shadow_params, params):
                    pc = p.detach().contiguous()
                    L.check(lib.dpb_ema_update(L.ptr(s), L.ptr(pc), s.numel(), float(omd), L.ptr(omd_dev),
                                               L.current_stream(s.device)))

    def copy_to(self, parameters):
        """ema.py:52-63 (in-place ``copy_`` on the parameter itself, so that the model notices the new weights)."""
        with torch.no_grad():
            for s, p in zip(self.shadow_params, [p for p in parameters if p.requires_grad]):
                p.copy_(s)

    def store(self, parameters):
        self.collected_params = [p.detach().clone() for p in parameters]

    def restore(self, parameters):
        with torch.no_grad():
            for c, p in zip(self.collected_params, parameters):
                p.copy_(c)

    def state_dict(self):
        return dict(decay=self.decay, num_updates=self.num_updates, shadow_params=self.shadow_params)

    def load_state_dict(self, state_dict):
        self.decay = state_dict['decay']
        self.num_updates = state_dict['num_updates']
        vals = list(state_dict['shadow_params'])
        self._alloc([v.to(self._flat.device) for v in vals], vals)
