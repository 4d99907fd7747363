"""Autograd bridge shared by the drop-in modules: forward/backward of a whole network through a UNetEngine."""
import threading

import torch


class EngineFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        engine = module._engine_for(x.device)
        names = module._param_names
        tensors = dict(zip(names, params))
        tensors.update(dict(module.named_buffers()))          # BatchNorm running statistics (updated in place)
        needs = any(ctx.needs_input_grad[2:])
        trainable = {n for n, need in zip(names, ctx.needs_input_grad[2:]) if need}
        logits = engine.forward(x, tensors, save=needs, training=module.training, trainable=trainable)
        if needs:
            engine._generation = getattr(engine, "_generation", 0) + 1
            ctx.generation = engine._generation
            ctx.engine = engine
            ctx.names = names
            ctx.params = params
            ctx.tensors = tensors
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        engine = ctx.engine
        if engine._generation != ctx.generation:
            raise RuntimeError("backward: the saved activations were overwritten by a later training forward")
        grads = {n: torch.empty_like(p) for n, p, need in zip(ctx.names, ctx.params, ctx.needs_input_grad[2:]) if need}
        engine.backward(dlogits, ctx.tensors, grads)
        return (None, None) + tuple(grads.get(n) for n in ctx.names)


class EngineModuleMixin:
    """Gives an nn.Module (a pure parameter container) a forward() that runs on the CUDA engine.  One engine per
    device: nn.DataParallel replicas share this object's dict, each device thread gets its own engine."""

    def _init_engine_state(self):
        self._engines = {}
        self._engine_lock = threading.Lock()

    def _make_engine(self, device):
        raise NotImplementedError

    def _engine_for(self, device):
        key = (device.type, device.index)
        with self._engine_lock:
            eng = self._engines.get(key)
            if eng is None:
                eng = self._make_engine(device)
                self._engines[key] = eng
            return eng

    @property
    def _param_names(self):
        return [n for n, _ in self.named_parameters()]

    def _engine_forward(self, inputs):
        if not inputs.is_cuda:
            raise RuntimeError(f"unet_pytorch_b200.{type(self).__name__} runs on a B200 only: move the module and inputs "
                               "to CUDA (there is no CPU fallback)")
        params = [p for _, p in self.named_parameters()]
        return EngineFunction.apply(self, inputs, *params)
