"""Minimal SQLite document-metadata table (python/fast_plaid/filtering/filtering.py).

Out of scope for the search hot path: kept only so that
``FastPlaid.create(metadata=...)`` and ``where(...) -> subset`` keep working.  Rows are keyed by
``_subset_`` = document id, exactly like the reference's METADATA table (filtering.py:179-199).
"""

from __future__ import annotations

import json
import os
import re
import sqlite3
from datetime import date, datetime
from typing import Any

_IDENT = re.compile(r"^[a-zA-Z_][a-zA-Z0-9_]*$")


def _check_column(name: str) -> None:
    """Column names are interpolated into SQL: only plain identifiers (filtering.py:10-12, :159-165)."""
    if not isinstance(name, str) or _IDENT.match(name) is None:
        raise ValueError(
            f"Invalid column name '{name}'. Column names must start with a letter or underscore, followed by "
            "letters, digits, or underscores, and cannot contain spaces or special characters."
        )


def _sql_type(value: Any) -> str:
    """filtering.py:15-25"""
    if isinstance(value, bool) or isinstance(value, int):
        return "INTEGER"
    if isinstance(value, float):
        return "REAL"
    if isinstance(value, (datetime, date, str)):
        return "TEXT"
    return "BLOB"


def _db(index: str) -> str:
    return os.path.join(index, "metadata.db")


def _columns(conn: sqlite3.Connection) -> list[str]:
    return [r[1] for r in conn.execute("PRAGMA table_info(METADATA)")]


def _insert(conn: sqlite3.Connection, start: int, metadata: list[dict[str, Any]]) -> None:
    cols = _columns(conn)
    for row in metadata:
        for k in row:
            _check_column(k)
            if k not in cols:
                first = next((r[k] for r in metadata if k in r and r[k] is not None), None)
                conn.execute(f'ALTER TABLE METADATA ADD COLUMN "{k}" {_sql_type(first)}')
                cols.append(k)
    for i, row in enumerate(metadata):
        keys = list(row.keys())
        vals = [json.dumps(v) if isinstance(v, (list, dict)) else (v.isoformat() if hasattr(v, "isoformat") else v)
                for v in row.values()]
        names = ", ".join(['"_subset_"'] + [f'"{k}"' for k in keys])
        marks = ", ".join(["?"] * (len(keys) + 1))
        conn.execute(f"INSERT INTO METADATA ({names}) VALUES ({marks})", [start + i] + vals)


def create(index: str, metadata: list[dict[str, Any]]) -> None:
    os.makedirs(index, exist_ok=True)
    if os.path.exists(_db(index)):
        os.remove(_db(index))
    with sqlite3.connect(_db(index)) as conn:
        conn.execute('CREATE TABLE METADATA ("_subset_" INTEGER PRIMARY KEY)')
        _insert(conn, 0, metadata)


def update(index: str, metadata: list[dict[str, Any]]) -> None:
    if not os.path.exists(_db(index)):
        return create(index, metadata)
    with sqlite3.connect(_db(index)) as conn:
        n = conn.execute('SELECT COALESCE(MAX("_subset_"), -1) + 1 FROM METADATA').fetchone()[0]
        _insert(conn, n, metadata)


def delete(index: str, subset: list[int]) -> None:
    if not os.path.exists(_db(index)):
        return
    with sqlite3.connect(_db(index)) as conn:
        conn.executemany('DELETE FROM METADATA WHERE "_subset_" = ?', [(int(i),) for i in subset])
        rows = [r[0] for r in conn.execute('SELECT "_subset_" FROM METADATA ORDER BY "_subset_"')]
        for new, old in enumerate(rows):  # renumber like the index does
            if new != old:
                conn.execute('UPDATE METADATA SET "_subset_" = ? WHERE "_subset_" = ?', (new, old))


def where(index: str, condition: str, parameters: tuple = ()) -> list[int]:
    if not os.path.exists(_db(index)):
        return []
    with sqlite3.connect(_db(index)) as conn:
        return [r[0] for r in conn.execute(f'SELECT "_subset_" FROM METADATA WHERE {condition} ORDER BY "_subset_"', parameters)]


def get(index: str, condition: str | None = None, parameters: tuple = (), subset: list[int] | None = None) -> list[dict]:
    if not os.path.exists(_db(index)):
        return []
    with sqlite3.connect(_db(index)) as conn:
        conn.row_factory = sqlite3.Row
        q = "SELECT * FROM METADATA"
        if condition:
            q += f" WHERE {condition}"
        rows = [dict(r) for r in conn.execute(q + ' ORDER BY "_subset_"', parameters)]
    if subset is not None:
        by_id = {r["_subset_"]: r for r in rows}
        rows = [by_id[i] for i in subset if i in by_id]
    return rows
