"""Command line for the compare / diagnostics path -- the hot-path commands of the reference's
``mcmc-ref`` CLI (cli.py:55-97 stats, :130-155 diagnostics, :158-163 info, :166-183 compare,
:186-200 convert) on the ``b200`` backend.  Same options, output formats and exit codes
(compare: 0 passed / 2 failed).  Run as ``python -m mcmc_ref_b200.cli``."""

from __future__ import annotations

import json
from pathlib import Path

import click

from . import convert as convert_mod
from . import reference
from .store import PackagedStore

_FORMATS3 = click.Choice(["table", "csv", "json"], case_sensitive=False)


def _headers(stats: dict[str, dict[str, float]]) -> list[str]:
    return sorted({k for metrics in stats.values() for k in metrics})


def _echo_table(stats: dict[str, dict[str, float]]) -> None:
    headers = ["param", *_headers(stats)]
    widths = [max(len(h), 6) for h in headers]
    click.echo(" ".join(h.ljust(w) for h, w in zip(headers, widths)))
    for param, metrics in stats.items():
        cells = [param, *(f"{metrics.get(h, float('nan')):.6g}" for h in headers[1:])]
        click.echo(" ".join(c.ljust(w) for c, w in zip(cells, widths)))


def _echo_csv(stats: dict[str, dict[str, float]], headers: list[str] | None = None, blank: str = "") -> None:
    headers = headers or _headers(stats)
    click.echo(",".join(["param", *headers]))
    for param, metrics in stats.items():
        click.echo(",".join([param, *(str(metrics.get(h, blank)) for h in headers)]))


@click.group()
def main() -> None:
    """mcmc-ref hot path on B200."""


@main.command("list")
@click.option("--format", "format_", type=click.Choice(["table", "json"], case_sensitive=False), default="table")
def list_cmd(format_: str) -> None:
    models = PackagedStore().list_models()
    click.echo(json.dumps(models, indent=2) if format_ == "json" else "\n".join(models))


@main.command("stats")
@click.argument("model")
@click.option("--params", default=None, help="Comma-separated parameter list")
@click.option("--format", "format_", type=_FORMATS3, default="table")
@click.option("--backend", type=click.Choice(["b200"], case_sensitive=False), default="b200")
@click.option("--quantile-mode", type=click.Choice(["exact"], case_sensitive=False), default="exact")
@click.option("--include-diagnostics", is_flag=True, help="Include rhat/ess metrics")
def stats_cmd(model: str, params: str | None, format_: str, backend: str, quantile_mode: str,
              include_diagnostics: bool) -> None:
    wanted = params.split(",") if params else None
    stats = reference.stats(model, params=wanted, backend=backend, quantile_mode=quantile_mode)
    if include_diagnostics:
        for param, metrics in reference.diagnostics_for_model(model, params=wanted).items():
            stats.setdefault(param, {}).update(metrics)
    if format_ == "json":
        click.echo(json.dumps(stats, indent=2, sort_keys=True))
    elif format_ == "csv":
        _echo_csv(stats)
    else:
        _echo_table(stats)


@main.command("diagnostics")
@click.argument("model")
@click.option("--format", "format_", type=_FORMATS3, default="table")
@click.option("--recompute", is_flag=True, help="Ignore the diagnostics cached in meta.json and run the device pass")
def diagnostics_cmd(model: str, format_: str, recompute: bool) -> None:
    diag = reference.diagnostics_for_model(model, recompute=recompute)
    if format_ == "json":
        click.echo(json.dumps(diag, indent=2, sort_keys=True))
    elif format_ == "csv":
        _echo_csv(diag, ["rhat", "ess_bulk", "ess_tail"], blank="None")
    else:
        _echo_table(diag)


@main.command("info")
@click.argument("model")
def info_cmd(model: str) -> None:
    click.echo(json.dumps(PackagedStore().read_meta(model), indent=2, sort_keys=True))


def _read_actual_csv(path: Path) -> dict:
    """Every non chain/draw column, all chains pooled (cli.py:294-297)."""
    import pyarrow.csv as pacsv

    table = pacsv.read_csv(path)
    return {c: table.column(c).to_numpy(zero_copy_only=False) for c in table.column_names if c not in {"chain", "draw"}}


@main.command("compare")
@click.argument("model")
@click.option("--actual", "actual_path", type=click.Path(path_type=Path), required=True)
@click.option("--tolerance", default=0.15, type=float)
@click.option("--format", "format_", type=click.Choice(["table", "json"], case_sensitive=False), default="table")
def compare_cmd(model: str, actual_path: Path, tolerance: float, format_: str) -> None:
    result = reference.compare(model, actual=_read_actual_csv(actual_path), tolerance=tolerance)
    if format_ == "json":
        details = {p: {k: vars(v) for k, v in metrics.items()} for p, metrics in result.details.items()}
        click.echo(json.dumps({"passed": result.passed, "failures": result.failures, "details": details},
                              indent=2, sort_keys=True))
    else:
        click.echo("passed" if result.passed else "failed")
        for failure in result.failures:
            click.echo(f"- {failure}")
    raise SystemExit(0 if result.passed else 2)


@main.command("log-density")
@click.argument("model")
@click.option("--draws", "draws_path", type=click.Path(path_type=Path), default=None,
              help="CSV of candidate draws (columns named like the packaged draws); default: the packaged reference draws")
@click.option("--format", "format_", type=click.Choice(["summary", "csv", "json"], case_sensitive=False), default="summary")
def log_density_cmd(model: str, draws_path: Path | None, format_: str) -> None:
    """EXTENSION (the reference has no such command): likelihood-block log-density of a packaged GLM-style
    model over draws, evaluated on the tensor cores (models.packaged_lp)."""
    draws = _read_actual_csv(draws_path) if draws_path is not None else None
    lp = reference.log_density(model, draws=draws).cpu().numpy()
    if format_ == "csv":
        click.echo("lp")
        for v in lp:
            click.echo(repr(float(v)))
    elif format_ == "json":
        click.echo(json.dumps([float(v) for v in lp]))
    else:
        click.echo(f"draws {lp.size} mean {lp.mean():.6g} sd {lp.std():.6g} min {lp.min():.6g} max {lp.max():.6g}")


@main.command("models")
def models_cmd() -> None:
    """Packaged models the log-density registry restates (family and Stan source lines)."""
    from .models import MODELS

    for name in sorted(MODELS):
        click.echo(f"{name}\t{MODELS[name].family}\t{MODELS[name].source}")


@main.command("convert")
@click.argument("input_path", type=click.Path(path_type=Path))
@click.option("--name", required=True)
@click.option("--force", is_flag=True)
def convert_cmd(input_path: Path, name: str, force: bool) -> None:
    from .store import _default_local_root

    root = _default_local_root()
    draws_dir, meta_dir = root / "draws", root / "meta"
    draws_dir.mkdir(parents=True, exist_ok=True)
    meta_dir.mkdir(parents=True, exist_ok=True)
    convert_mod.convert_file(input_path, name=name, out_draws_dir=draws_dir, out_meta_dir=meta_dir, force=force)
    click.echo(f"converted {name} -> {draws_dir}")


if __name__ == "__main__":
    main()
