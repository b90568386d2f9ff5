"""Attack base class -- mirrors PointNet/attacks/torchattacks/attack.py:4-195 (eval-mode switch,
``__call__``, attack-mode / return-type plumbing).  The scripts of the reference never call
``set_attack_mode``, so ``_targeted`` stays +1; the setter is kept with
the reference's semantics."""
from __future__ import annotations

import numpy as np
import torch


class Attack(object):
    def __init__(self, name, model):
        self.attack = name
        self.model = model
        self.model_name = str(model).split("(")[0]
        self.training = model.training
        self.device = next(model.parameters()).device
        self._targeted = 1
        self._attack_mode = "default"
        self._return_type = "float"
        self._target_map_function = lambda images, labels: labels

    def forward(self, *input):
        raise NotImplementedError

    def set_attack_mode(self, mode, target_map_function=None):
        if self._attack_mode == "only_default":
            raise ValueError("Changing attack mode is not supported in this attack method.")
        if mode == "targeted" and target_map_function is None:
            raise ValueError("Please give a target_map_function, e.g., lambda images, labels:(labels+1)%10.")
        if mode == "default":
            self._attack_mode, self._targeted = "default", 1
        elif mode == "targeted":
            self._attack_mode, self._targeted = "targeted", -1
            self._target_map_function = target_map_function
        elif mode == "least_likely":
            self._attack_mode, self._targeted = "least_likely", -1
        else:
            raise ValueError(mode + " is not a valid mode. [Options : default, targeted, least_likely]")

    def set_return_type(self, type):
        if type == "float":
            self._return_type = "float"
        elif type == "int":
            self._return_type = "int"
        else:
            raise ValueError(type + " is not a valid type. [Options : float, int]")

    def _to_uint(self, images):
        return (images * 255).type(torch.uint8)

    def _switch_model(self):
        # attack.py:52-56.  The module-tree walk of train() / eval() is ~0.3 ms of host time per call during which the GPU
        # has nothing to do, so it is skipped when the root already is in the wanted mode (the engine reads only that flag).
        if self.model.training != bool(self.training):
            self.model.train(bool(self.training))

    def __str__(self):
        info = {k: v for k, v in self.__dict__.items() if k[0] != "_" and k not in ("model", "attack")}
        info["attack_mode"] = "default" if self._attack_mode == "only_default" else self._attack_mode
        info["return_type"] = self._return_type
        return self.attack + "(" + ", ".join("{}={}".format(k, v) for k, v in info.items()) + ")"

    def __call__(self, *input, **kwargs):
        # run on the images' device (the C library launches on the current device)
        if input and torch.is_tensor(input[0]) and input[0].is_cuda and input[0].device.index != torch.cuda.current_device():
            with torch.cuda.device(input[0].device):
                return self.__call__(*input, **kwargs)
        if self.model.training:
            self.model.eval()
        images = self.forward(*input, **kwargs)
        self._switch_model()
        if self._return_type == "int":
            images = self._to_uint(images)
        return images

    # ---- helpers shared by the four attacks ------------------------------------------------------
    def _engine(self, images):
        eng_of = getattr(self.model, "engine", None)
        if eng_of is None:
            raise TypeError("pointsecguard_b200 attacks drive pointsecguard_b200 models (get_model of "
                            "models/pointnet2_sem_seg{,_msg}.py); there is no generic autograd fallback")
        if not images.is_cuda:
            raise RuntimeError("pointsecguard_b200 attacks need CUDA tensors; there is no CPU fallback")
        return eng_of(images.device)

    @staticmethod
    def _labels_i32(labels, device):
        """labels arrive as a numpy float64 [B,N] array in the scripts (nontarget.py:25, Q15)."""
        if torch.is_tensor(labels):
            return labels.to(device=device, dtype=torch.int32).contiguous()
        return torch.as_tensor(np.asarray(labels).astype(np.int32), device=device).contiguous()

    @staticmethod
    def _mask_u8(mask, B, N, device):
        """mask: numpy / torch bool [N] (reference, B == 1) or [B,N] (per-block generalisation)."""
        m = torch.as_tensor(np.asarray(mask)) if not torch.is_tensor(mask) else mask
        m = m.to(device=device, dtype=torch.bool)
        if m.dim() == 1:
            m = m.view(1, N).expand(B, N)
        return m.to(torch.uint8).contiguous()
