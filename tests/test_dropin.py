"""The zero-edit recipe of INTEGRATION.md section 1, executed: install_as_reference_module() must replace ONLY the two
leaf modules of lib.csrc.ransac_voting and leave every other `lib.*` import of a clean-pvnet checkout working
(lib/networks/pvnet/resnet18.py:5-6, lib/evaluators/linemod/pvnet.py:13,20, run.py:1)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PVNET_REFERENCE", "/root/reference")


def _run(code, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], cwd=cwd, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    return r.stdout


def _fake_checkout(tmp_path, with_leaf_package):
    """A miniature clean-pvnet tree: regular package `lib`, namespace package `lib/csrc` (the reference has no
    lib/csrc/__init__.py), sibling sub-packages that must stay importable."""
    files = {
        "lib/__init__.py": "",
        "lib/config/__init__.py": "from .config import cfg\n",
        "lib/config/config.py": "class _C: pass\ncfg = _C()\ncfg.marker = 'real lib.config'\n",
        "lib/networks/__init__.py": "from .pvnet import resnet18\n",
        "lib/networks/pvnet/__init__.py": "",
        "lib/networks/pvnet/resnet18.py":
            "from lib.csrc.ransac_voting.ransac_voting_gpu import ransac_voting_layer, ransac_voting_layer_v3, "
            "estimate_voting_distribution_with_mean\nfrom lib.config import cfg\n",
        "lib/csrc/nn/nn_utils.py": "MARK = 'real nn_utils'\n",
        "lib/csrc/uncertainty_pnp/un_pnp_utils.py": "MARK = 'real un_pnp_utils'\n",
        "lib/utils/pvnet/pvnet_pose_utils.py": "MARK = 'real pose utils'\n",
        "lib/utils/__init__.py": "",
        "lib/utils/pvnet/__init__.py": "",
    }
    if with_leaf_package:
        # what the checkout really holds: the reference operator whose import of the (unbuilt) pybind module would fail
        files["lib/csrc/ransac_voting/ransac_voting_gpu.py"] = (
            "import lib.csrc.ransac_voting.ransac_voting as ransac_voting\nraise RuntimeError('reference operator imported')\n")
    for rel, text in files.items():
        p = tmp_path / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(text)
    return str(tmp_path)


CHECK = """
    import sys
    sys.path.insert(0, {tree!r})
    import clean_pvnet_b200
    clean_pvnet_b200.install_as_reference_module()          # INTEGRATION.md section 1, verbatim order
    import lib.networks                                      # -> resnet18.py:5-6 inside
    from lib.networks.pvnet import resnet18
    assert resnet18.ransac_voting_layer_v3 is clean_pvnet_b200.ransac_voting_layer_v3
    assert resnet18.ransac_voting_layer is clean_pvnet_b200.ransac_voting_layer
    assert resnet18.estimate_voting_distribution_with_mean is clean_pvnet_b200.estimate_voting_distribution_with_mean
    assert resnet18.cfg.marker == 'real lib.config'
    from lib.csrc.nn import nn_utils
    from lib.csrc.uncertainty_pnp import un_pnp_utils
    from lib.utils.pvnet import pvnet_pose_utils
    assert nn_utils.MARK == 'real nn_utils' and un_pnp_utils.MARK == 'real un_pnp_utils'
    import lib, lib.csrc
    assert not getattr(lib, '__pvb_stand_in__', False) and not getattr(lib.csrc, '__pvb_stand_in__', False)
    assert lib.__file__.startswith({tree!r})
    import lib.csrc.ransac_voting.ransac_voting as ext
    assert ext is clean_pvnet_b200.ransac_voting
    clean_pvnet_b200.install_as_reference_module()          # idempotent
    print('ok')
"""


@pytest.mark.parametrize("with_leaf_package", [True, False])
def test_install_inside_a_checkout_keeps_every_other_lib_import(tmp_path, with_leaf_package):
    tree = _fake_checkout(tmp_path, with_leaf_package)
    assert "ok" in _run(CHECK.format(tree=tree), cwd=tree)


def test_install_after_lib_was_already_imported(tmp_path):
    tree = _fake_checkout(tmp_path, True)
    code = """
        import sys
        sys.path.insert(0, {tree!r})
        import lib.config                                    # e.g. `from lib.config import cfg` at the top of run.py
        import clean_pvnet_b200
        clean_pvnet_b200.install_as_reference_module()
        import lib.networks
        from lib.csrc.nn import nn_utils
        print('ok')
    """
    assert "ok" in _run(code.format(tree=tree), cwd=tree)


def test_install_without_a_checkout_uses_stand_ins(tmp_path):
    code = """
        import clean_pvnet_b200
        clean_pvnet_b200.install_as_reference_module()
        from lib.csrc.ransac_voting.ransac_voting_gpu import ransac_voting_layer_v3
        import lib
        assert lib.__pvb_stand_in__
        print('ok')
    """
    assert "ok" in _run(code, cwd=str(tmp_path))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "lib", "networks", "pvnet")), reason="reference checkout absent")
def test_the_real_resnet18_import_line_resolves_to_this_package():
    """lib/networks/pvnet/resnet18.py:5 of the UNMODIFIED reference, imported from where it lies.  Two modules the
    reference needs do not exist in this container and are stubbed: `imp` (removed in Python 3.12;
    lib/networks/make_network.py:2) and `lib.config` (needs yacs + open3d, lib/config/config.py:1-5)."""
    code = """
        import sys, types
        sys.path.insert(0, {ref!r})
        sys.modules['imp'] = types.ModuleType('imp')
        cfgmod = types.ModuleType('lib.config'); cfgmod.cfg = types.SimpleNamespace(); sys.modules['lib.config'] = cfgmod
        import clean_pvnet_b200
        clean_pvnet_b200.install_as_reference_module()
        import lib
        assert lib.__file__.startswith({ref!r}), lib.__file__
        from lib.networks.pvnet import resnet18                # runs resnet18.py:1-6 of the reference
        assert resnet18.ransac_voting_layer_v3 is clean_pvnet_b200.ransac_voting_layer_v3
        assert resnet18.estimate_voting_distribution_with_mean is clean_pvnet_b200.estimate_voting_distribution_with_mean
        assert callable(resnet18.Resnet18.decode_keypoint)     # the caller of the operator (resnet18.py:65-76)
        assert resnet18.Resnet18.decode_keypoint.__globals__['ransac_voting_layer_v3'] is clean_pvnet_b200.ransac_voting_layer_v3
        import importlib.util
        for name in ('lib.csrc.nn.nn_utils', 'lib.csrc.uncertainty_pnp.un_pnp_utils', 'lib.utils.pvnet.pvnet_pose_utils',
                     'lib.networks.ct_pvnet'):
            assert importlib.util.find_spec(name) is not None, name
        print('ok')
    """
    assert "ok" in _run(code.format(ref=REF), cwd=REF)
