"""The Attack base class keeps the reference's plumbing (PointNet/attacks/torchattacks/attack.py:4-195): attack-mode and
return-type setters with their error behaviour, the string form, the train/eval switch around a call.  CPU only: nothing
here launches a kernel (the attacks themselves refuse CPU tensors -- there is no fallback)."""
import pytest
import torch

from pointsecguard_b200 import torchattacks
from pointsecguard_b200.models.pointnet2_sem_seg import get_model
from pointsecguard_b200.torchattacks.attack import Attack


def test_attack_mode_setter_matches_reference_semantics():
    atk = torchattacks.NB_attack(get_model(13), eps=0.1, alpha=0.05, iters=3)
    assert atk._attack_mode == "default" and atk._targeted == 1
    with pytest.raises(ValueError):
        atk.set_attack_mode("targeted")                                   # needs a target_map_function (attack.py:72-74)
    f = lambda images, labels: (labels + 1) % 13
    atk.set_attack_mode("targeted", f)
    assert atk._attack_mode == "targeted" and atk._targeted == -1 and atk._target_map_function is f
    atk.set_attack_mode("least_likely")
    assert atk._attack_mode == "least_likely" and atk._targeted == -1
    atk.set_attack_mode("default")
    assert atk._attack_mode == "default" and atk._targeted == 1
    with pytest.raises(ValueError):
        atk.set_attack_mode("something_else")
    atk._attack_mode = "only_default"
    with pytest.raises(ValueError):
        atk.set_attack_mode("default")


def test_return_type_setter_and_uint_conversion():
    atk = torchattacks.tar_NB_attack(get_model(13), eps=0.5, alpha=0.1, iters=2, target=7, mask=None)
    atk.set_return_type("int")
    assert atk._return_type == "int"
    x = torch.tensor([0.0, 0.5, 1.0])
    assert atk._to_uint(x).dtype == torch.uint8 and atk._to_uint(x).tolist() == [0, 127, 255]
    atk.set_return_type("float")
    assert atk._return_type == "float"
    with pytest.raises(ValueError):
        atk.set_return_type("double")


def test_string_form_lists_the_public_fields():
    atk = torchattacks.NB_attack(get_model(13), eps=0.1, alpha=0.05, iters=3)
    s = str(atk)
    assert s.startswith("NB_attack(") and "eps=0.1" in s and "iters=3" in s and "attack_mode=default" in s and "return_type=float" in s


def test_cpu_tensors_are_refused_and_the_model_mode_is_restored():
    m = get_model(13).train()
    atk = torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=1)
    assert atk.training is True
    with pytest.raises(RuntimeError):
        atk(torch.zeros(1, 9, 64), torch.zeros(1, 64).numpy())
    assert isinstance(atk, Attack)
