"""Host-side mirror of the NeMo-0.10 neural-module boundary the hot path sits behind.

Only what `infer.py` touches: typed ports, the symbolic
``module(**NmTensors)`` call that builds the DAG, ``force_pt=True`` execution,
``restore_from``/``save_to``, and ``NeuralModuleFactory.infer``.  Names,
argument meaning and error behaviour follow the reference:

  * NeuralType / NmTensor          nemo/core/neural_types/neural_type.py:29-237
  * NeuralModule.__call__          nemo/core/neural_modules.py:423-523
  * TrainableNM / NonTrainableNM / DataLayerNM
                                   nemo/backends/pytorch/nm.py:13-129, 132-184, 187-320
  * NeuralModuleFactory(.infer)    nemo/core/neural_factory.py:251-415, 623-671
  * DAG execution                  nemo/backends/pytorch/actions.py:380-442, 639-821

Everything numerical is delegated to the CUDA library; this file is plumbing.
"""
from __future__ import annotations

import enum
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn


# ----------------------------------------------------------------------------- neural types
class ElementType:
    """nemo/core/neural_types/elements.py - only the hierarchy the ASR ports use."""

    def __init__(self, **params):
        self.params = params

    def __repr__(self):
        return type(self).__name__


class VoidType(ElementType): pass
class ChannelType(ElementType): pass
class AudioSignal(ElementType):
    def __init__(self, freq: int = 16000):
        super().__init__(freq=freq)
class SpectrogramType(ChannelType): pass
class MelSpectrogramType(SpectrogramType): pass
class AcousticEncodedRepresentation(ChannelType): pass
class LengthsType(ElementType): pass
class LogprobsType(ElementType): pass
class PredictionsType(ElementType): pass


class NeuralTypeComparisonResult(enum.Enum):
    SAME = 0
    LESS = 1
    GREATER = 2
    DIM_INCOMPATIBLE = 3
    INCOMPATIBLE = 6


class NeuralTypeError(Exception): pass
class NeuralPortNameMismatchError(NeuralTypeError): pass
class NeuralPortNmTensorMismatchError(NeuralTypeError): pass


class NeuralType:
    def __init__(self, axes: Optional[Tuple] = None, elements_type: ElementType = None, optional=False):
        self.axes = tuple(axes) if axes is not None else None
        self.elements_type = elements_type if elements_type is not None else VoidType()
        self.optional = optional

    def compare(self, second: "NeuralType") -> NeuralTypeComparisonResult:
        """neural_type.py:77 - SAME / LESS (second is a subtype) are accepted by __call__."""
        if isinstance(self.elements_type, VoidType) and self.axes is None:
            return NeuralTypeComparisonResult.SAME
        if self.axes is not None and second.axes is not None and len(self.axes) != len(second.axes):
            return NeuralTypeComparisonResult.DIM_INCOMPATIBLE
        a, b = type(self.elements_type), type(second.elements_type)
        if a is b:
            if self.elements_type.params != second.elements_type.params:
                return NeuralTypeComparisonResult.INCOMPATIBLE
            return NeuralTypeComparisonResult.SAME
        if issubclass(b, a):
            return NeuralTypeComparisonResult.GREATER
        if issubclass(a, b):
            return NeuralTypeComparisonResult.LESS
        return NeuralTypeComparisonResult.INCOMPATIBLE

    def __repr__(self):
        return f"NeuralType({self.axes}, {self.elements_type!r})"


class NmTensor(NeuralType):
    """Symbolic edge of the DAG (neural_type.py:185-237)."""

    def __init__(self, producer, producer_args, name, ntype: NeuralType):
        super().__init__(ntype.axes, ntype.elements_type, ntype.optional)
        self._producer = producer
        self._producer_args = producer_args
        self._name = name

    @property
    def producer(self): return self._producer
    @property
    def producer_args(self): return self._producer_args
    @property
    def name(self): return self._name
    @property
    def unique_name(self): return f"{self._name}~~~{id(self._producer)}"


# ----------------------------------------------------------------------------- factory
class DeviceType(enum.Enum):
    GPU = 1
    CPU = 2
    AllGpu = 3


class NeuralModuleFactory:
    """nemo/core/neural_factory.py:251-415.  Inference only: no trainer, no ExpManager."""

    _DEFAULT = None

    def __init__(self, backend=None, local_rank=None, optimization_level=None, placement=None,
                 cudnn_benchmark=False, random_seed=None, **_ignored):
        self._local_rank = local_rank
        self._world_size = 1
        if placement is None:
            placement = DeviceType.AllGpu if local_rank is not None else DeviceType.GPU
        self._placement = placement
        if placement == DeviceType.AllGpu and torch.distributed.is_available() and torch.distributed.is_initialized():
            self._world_size = torch.distributed.get_world_size()
        if random_seed is not None:
            torch.manual_seed(random_seed)
        NeuralModuleFactory._DEFAULT = self

    @classmethod
    def get_default_factory(cls):
        return cls._DEFAULT

    @classmethod
    def set_default_factory(cls, factory):
        cls._DEFAULT = factory

    @property
    def placement(self): return self._placement
    @property
    def world_size(self): return self._world_size
    @property
    def local_rank(self): return self._local_rank

    # ---- actions.py:1423-1488 (infer) + :639-821 (_infer) + :380-442 (forward pass)
    def infer(self, tensors: List[NmTensor], checkpoint_dir=None, ckpt_pattern="", verbose=True,
              cache=False, use_cache=False, offload_to_cpu=True, modules_to_restore=None):
        data_layers = set()

        def find_dl(t: NmTensor):
            p = t.producer
            if isinstance(p, DataLayerNM):
                data_layers.add(p)
            for a in (t.producer_args or {}).values():
                find_dl(a)

        for t in tensors:
            find_dl(t)
        if len(data_layers) != 1:
            raise ValueError(f"There should be exactly one DataLayer in the call chain, found {len(data_layers)}")
        dl = next(iter(data_layers))
        results = [[] for _ in tensors]
        with torch.no_grad():
            for batch in dl.data_iterator:
                if not isinstance(batch, (tuple, list)):
                    batch = (batch,)
                values: Dict[str, object] = {}
                for (port, _), v in zip(dl.output_ports.items(), batch):
                    if isinstance(v, torch.Tensor):
                        v = v.to(dl._device)
                    values[f"{port}~~~{id(dl)}"] = v

                def evaluate(t: NmTensor):
                    if t.unique_name in values:
                        return values[t.unique_name]
                    m = t.producer
                    call = {k: evaluate(a) for k, a in t.producer_args.items()}
                    if isinstance(m, nn.Module):
                        m.eval()                                   # actions.py:414-415
                    out = m(force_pt=True, **call)                # actions.py:428
                    if not isinstance(out, (tuple, list)):
                        out = (out,)
                    for (port, _), v in zip(m.output_ports.items(), out):
                        values[f"{port}~~~{id(m)}"] = v
                    return values[t.unique_name]

                for i, t in enumerate(tensors):
                    v = evaluate(t)
                    if offload_to_cpu and isinstance(v, torch.Tensor):
                        v = v.cpu()                                # actions.py:808-812
                    results[i].append(v)
        return results


def _get_device(placement):
    if placement in (DeviceType.GPU, DeviceType.AllGpu):
        return torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
    return torch.device("cpu")


# ----------------------------------------------------------------------------- modules
class NeuralModule:
    """nemo/core/neural_modules.py:52-128 (ctor) and :423-523 (symbolic __call__)."""

    def __init__(self):
        self._factory = NeuralModuleFactory.get_default_factory()
        if self._factory is None:
            # neural_modules.py:66-74 creates a default factory with a warning
            self._factory = NeuralModuleFactory()
        self._placement = self._factory.placement
        self._opt_level = None

    @property
    def placement(self): return self._placement
    @property
    def factory(self): return self._factory
    @property
    def input_ports(self) -> Dict[str, NeuralType]: return {}
    @property
    def output_ports(self) -> Dict[str, NeuralType]: return {}

    def _symbolic_call(self, **kwargs):
        in_ports = self.input_ports
        for name, t in kwargs.items():
            if name not in in_ports:
                raise NeuralPortNameMismatchError(f"Wrong input port name: {name}")
            if not isinstance(t, NmTensor):
                raise NeuralPortNmTensorMismatchError(f"Port {name} expects an NmTensor, got {type(t).__name__}")
            res = in_ports[name].compare(t)
            if res not in (NeuralTypeComparisonResult.SAME, NeuralTypeComparisonResult.GREATER):
                raise NeuralPortNmTensorMismatchError(
                    f"\n\nIn {type(self).__name__}. \nPort: {name} and a NmTensor it was fed are \n"
                    f"of incompatible neural types:\n\n{in_ports[name]} \n\n and \n\n{t}\n\nType comparison result: {res}")
        for name, t in in_ports.items():
            if name not in kwargs and not t.optional:
                raise NeuralPortNameMismatchError(f"Input port {name} is required but was not provided")
        outs = tuple(NmTensor(self, dict(kwargs), name, t) for name, t in self.output_ports.items())
        return outs[0] if len(outs) == 1 else outs


class TrainableNM(NeuralModule, nn.Module):
    """nemo/backends/pytorch/nm.py:13-129."""

    def __init__(self, pretrained_model_name=None):
        NeuralModule.__init__(self)
        nn.Module.__init__(self)
        self._device = _get_device(self.placement)
        self._pretrained_model_name = pretrained_model_name

    def __call__(self, *input, force_pt=False, **kwargs):
        if len(input) > 0 or force_pt:
            return nn.Module.__call__(self, *input, **kwargs)
        return self._symbolic_call(**kwargs)

    def get_weights(self):
        return {n: (p, p.requires_grad) for n, p in self.named_parameters()}

    def save_to(self, path):
        torch.save(self.state_dict(), path)

    def restore_from(self, path, local_rank=0):
        dev = f"cuda:{local_rank}" if self.placement == DeviceType.AllGpu else self._device
        self.load_state_dict(torch.load(path, map_location=dev))

    def freeze(self, weights=None):
        for n, p in self.named_parameters():
            if weights is None or n in weights:
                p.requires_grad = False

    def unfreeze(self, weights=None):
        for n, p in self.named_parameters():
            if weights is None or n in weights:
                p.requires_grad = True

    @property
    def num_weights(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)


class NonTrainableNM(NeuralModule):
    """nemo/backends/pytorch/nm.py:132-184 (not an nn.Module)."""

    def __init__(self):
        NeuralModule.__init__(self)
        self._device = _get_device(self.placement)

    def __call__(self, force_pt=False, *input, **kwargs):
        if len(input) > 0 or force_pt:
            with torch.no_grad():
                return self.forward(*input, **kwargs)
        return self._symbolic_call(**kwargs)

    def forward(self, *input):
        raise NotImplementedError

    def get_weights(self): return None
    def save_to(self, path): pass
    def restore_from(self, path): pass
    def freeze(self, weights=None): pass
    def unfreeze(self, weights=None): pass

    @property
    def num_weights(self): return 0


class DataLayerNM(NeuralModule):
    """nemo/backends/pytorch/nm.py:187-320: subclasses give __len__, dataset / data_iterator."""

    def __init__(self):
        NeuralModule.__init__(self)
        self._device = _get_device(self.placement)

    def __call__(self, force_pt=False, *input, **kwargs):
        return self._symbolic_call(**kwargs)

    @property
    def input_ports(self): return {}

    def get_weights(self): return None
    def save_to(self, path): pass
    def restore_from(self, path): pass

    @property
    def num_weights(self): return 0


# ----------------------------------------------------------------------------- reference backend
# Inside the reference tree the drop-in modules subclass the reference's OWN base classes (INTEGRATION.md section 1:
# "change one import line in asr.py").  VASR_NM_BACKEND=nemo does exactly that without editing a file: every name
# asr.py / pipeline.py import from this module is re-bound to the class of the same name in the vendored NeMo 0.10
# (`nemo` must be importable).  tests/test_boundary_reference.py runs infer.py's wiring this way against
# /root/reference.
import os as _os

if _os.environ.get("VASR_NM_BACKEND") == "nemo":
    from nemo.backends.pytorch.nm import DataLayerNM, NonTrainableNM, TrainableNM      # noqa: F401,E402
    from nemo.core import DeviceType, NeuralModule, NeuralModuleFactory                  # noqa: F401,E402
    from nemo.core.neural_types import (AcousticEncodedRepresentation, AudioSignal, ChannelType, LengthsType,  # noqa: F401,E402
                                        LogprobsType, MelSpectrogramType, NeuralPortNameMismatchError,
                                        NeuralPortNmTensorMismatchError, NeuralType, NeuralTypeError, NmTensor,
                                        PredictionsType, SpectrogramType, VoidType)
