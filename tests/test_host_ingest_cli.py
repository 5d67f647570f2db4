"""CPU tests of the host logic around the hot path: the convert gate's table helpers, the
chain geometry of the ingestion step, the CLI wiring that needs no device (reference
counterparts: tests/unit/test_convert.py, tests/unit/test_cli.py:24-39)."""

from __future__ import annotations

import json
import zipfile
from pathlib import Path

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest
from click.testing import CliRunner


@pytest.fixture(scope="module", autouse=True)
def _paths():
    import __graft_entry__ as g

    g._ensure_path()


def _write_draws_parquet(path: Path) -> None:
    # the reference's CLI fixture, tests/unit/test_cli.py:13-21
    pq.write_table(pa.table({
        "chain": pa.array([0, 0, 1, 1, 2, 2, 3, 3], type=pa.int32()),
        "draw": pa.array([0, 1, 0, 1, 0, 1, 0, 1], type=pa.int32()),
        "mu": pa.array([1.0, 2.0, 1.5, 2.5, 1.0, 2.0, 1.5, 2.5], type=pa.float64()),
    }), path)


def test_ensure_chain_draw_and_counts():
    from mcmc_ref_b200 import convert

    t = convert._ensure_chain_draw(pa.table({"mu": [1.0, 2.0, 3.0]}))
    assert t.column("chain").to_pylist() == [0, 0, 0] and t.column("draw").to_pylist() == [0, 1, 2]
    t = convert._ensure_chain_draw(pa.table({"chain": [0, 0, 1], "mu": [1.0, 2.0, 3.0]}))
    assert t.column("draw").to_pylist() == [0, 1, 2]
    t = convert._ensure_chain_draw(pa.table({"draw": [0, 1, 2], "mu": [1.0, 2.0, 3.0]}))
    assert t.column("chain").to_pylist() == [0, 0, 0]
    assert convert._count_chains_draws(pa.table({"chain": [0, 0, 0, 1, 1], "draw": [0, 1, 2, 0, 1]})) == (2, 2)
    assert convert._count_chains_draws(pa.table({"chain": pa.array([], type=pa.int32())})) == (0, 0)


def test_read_json_zip_layout(tmp_path):
    from mcmc_ref_b200 import convert

    payload = [{"b": [1.0, 2.0], "a": [3.0, 4.0]}, {"b": [5.0, 6.0], "a": [7.0, 8.0]}]
    path = tmp_path / "m.json.zip"
    with zipfile.ZipFile(path, "w") as zf:
        zf.writestr("m.json", json.dumps(payload))
    t = convert._read_input(path)
    assert t.column_names == ["chain", "draw", "a", "b"]
    assert t.column("chain").to_pylist() == [0, 0, 1, 1] and t.column("draw").to_pylist() == [0, 1, 0, 1]
    assert t.column("a").to_pylist() == [3.0, 4.0, 7.0, 8.0] and t.column("b").to_pylist() == [1.0, 2.0, 5.0, 6.0]
    with pytest.raises(ValueError, match="Unsupported input format"):
        convert._read_input(tmp_path / "x.txt")


def test_checks_and_enforce():
    from mcmc_ref_b200 import convert

    diag = {"a": {"rhat": 1.001, "ess_bulk": 900.0, "ess_tail": 800.0}, "b": {"rhat": 1.02, "ess_bulk": 300.0, "ess_tail": 1.0}}
    checks = convert._checks(10, 1000, diag)
    assert checks == {"ndraws_is_10k": True, "nchains_is_gte_4": True, "ess_above_400": False, "rhat_below_1_01": False}
    with pytest.raises(ValueError, match="quality checks failed: ess_above_400, rhat_below_1_01"):
        convert._enforce_checks(checks)
    convert._enforce_checks(convert._checks(10, 1000, {"a": diag["a"]}))


def test_chain_geometry_rejects_ragged():
    from mcmc_ref_b200 import ingest

    assert ingest.chain_geometry(np.array([0, 0, 1, 1, 2, 2])) == (3, 2)
    assert ingest.chain_geometry(np.array([], dtype=np.int64)) == (0, 0)
    with pytest.raises(ValueError, match="ragged chains"):
        ingest.chain_geometry(np.array([0, 0, 1]))


def test_cli_list_and_info_need_no_device(tmp_path, monkeypatch):
    from mcmc_ref_b200.cli import main

    (tmp_path / "draws").mkdir()
    (tmp_path / "meta").mkdir()
    _write_draws_parquet(tmp_path / "draws" / "example.draws.parquet")
    (tmp_path / "meta" / "example.meta.json").write_text(json.dumps({"model": "example"}))
    monkeypatch.setenv("MCMC_REF_LOCAL_ROOT", str(tmp_path))
    runner = CliRunner()
    res = runner.invoke(main, ["list", "--format", "json"])
    assert res.exit_code == 0 and "example" in res.output
    res = runner.invoke(main, ["info", "example"])
    assert res.exit_code == 0 and json.loads(res.output) == {"model": "example"}
    res = runner.invoke(main, ["stats", "example", "--backend", "numpy"])
    assert res.exit_code == 2  # click rejects backends this build does not register


def test_reference_api_surface_without_device(tmp_path):
    """list_models / draws / stan_data / model_code of the reference's Python API (reference.py:15-73;
    reference test: tests/unit/test_reference.py:13-84) through the packaged layout, no device needed."""
    from mcmc_ref_b200 import reference
    from mcmc_ref_b200.store import PackagedStore

    for sub in ("draws", "meta", "stan_data", "stan_models"):
        (tmp_path / sub).mkdir()
    _write_draws_parquet(tmp_path / "draws" / "m.draws.parquet")
    (tmp_path / "stan_data" / "m.data.json").write_text(json.dumps({"N": 2, "y": [1.0, 2.0]}))
    (tmp_path / "stan_models" / "m.stan").write_text("model { }\n")
    store = PackagedStore(packaged_root=tmp_path, local_root=tmp_path / "none")
    assert reference.list_models(store=store) == ["m"]
    assert reference.stan_data("m", store=store) == {"N": 2, "y": [1.0, 2.0]}
    assert reference.model_code("m", store=store).startswith("model")
    a = reference.draws("m", return_="numpy", store=store)
    assert a.shape == (8, 1) and a[1, 0] == 2.0
    t = reference.draws("m", params=["mu"], chains=[1, 3], return_="arrow", store=store).read_all()
    assert t.num_rows == 4 and set(t.column("chain").to_pylist()) == {1, 3}
    assert reference.draws("m", return_="list", store=store)[0] == {"chain": 0, "draw": 0, "mu": 1.0}
    with pytest.raises(ValueError, match="Unknown return type"):
        reference.draws("m", return_="frame", store=store)
    with pytest.raises(FileNotFoundError, match="draws not found for model"):
        reference.draws("absent", store=store)
    with pytest.raises(FileNotFoundError, match="stan data not found"):
        reference.stan_data("absent", store=store)
