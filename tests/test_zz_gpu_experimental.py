"""Opt-in GPU parity run of EXPERIMENTAL data layouts (not part of the default product path, not yet measured):
FBGPU_ARRAY_STRIPED=1 permutes array payloads at load time (featurebase_b200/csrc/stripe.h).  Every kernel must give
bit-identical results on permuted arrays, so the bodies of the regular parity tests are simply re-run with the switch on.
Skipped unless FBGPU_TEST_EXPERIMENTAL=1 (round 2 turns it on before measuring the layout)."""
import os

import pytest

from tests import test_gpu_parity as G

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.environ.get("FBGPU_TEST_EXPERIMENTAL"), reason="experimental layouts: set FBGPU_TEST_EXPERIMENTAL=1")]


@pytest.fixture
def striped(monkeypatch):
    monkeypatch.setenv("FBGPU_ARRAY_STRIPED", "1")      # read when a context is created


def test_striped_set_ops(striped):
    G.test_config1_single_shard_plumbing()
    G.test_container_combinations_table_on_gpu()
    G.test_mixed_encoding_pairs()
    G.test_union_intersect_count_config2_small()
    G.test_executor_goldens_and_edge_semantics()


@pytest.mark.parametrize("mode", [0, 1])
def test_striped_density_sweep(striped, mode):
    G.test_density_sweep_intersect_count(mode)


def test_striped_bsi_topk_groupby(striped):
    G.test_bsi_range_goldens_on_gpu()
    G.test_bsi_uniform_u32_config3_small()
    G.test_topk_topn_rowcounts()
    G.test_groupby_two_and_three_fields()


@pytest.mark.parametrize("env", ["FBGPU_FORCE_WORDPAR", "FBGPU_STAGED"])
def test_striped_alternative_kernels(striped, env, monkeypatch):
    G.test_alternative_eval_kernels(env, monkeypatch)       # FORCE_WORDPAR must be ignored for views that hold arrays
