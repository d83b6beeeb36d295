"""athinput parameter files -- host-side mirror of the reference's par.c interface.

The reference parses `<block>` / `name = value # comment` files and lets any value be overridden on the
command line as `block/name=value` (/root/reference/src/par.c:194 par_cmdline, :236-330 getters;
/root/reference/src/main.c:305-306).  The hot path only consumes the keys listed in SURVEY.md App. B; this
module gives the Python host driver the same lookup semantics (same function names, same "missing key is
fatal" behaviour, same `_def` variants) so decks written for the reference run unchanged.
"""
from __future__ import annotations

import re


class ParError(RuntimeError):
    """Raised where par.c would call ath_error() (missing block / key)."""


class ParFile:
    def __init__(self):
        self.blocks: dict[str, dict[str, str]] = {}

    # ---- par_open (par.c:155) ------------------------------------------------------------------
    @classmethod
    def par_open(cls, filename: str) -> "ParFile":
        with open(filename) as f:
            return cls.from_text(f.read())

    @classmethod
    def from_text(cls, text: str) -> "ParFile":
        p = cls()
        block = None
        for raw in text.splitlines():
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            m = re.match(r"<\s*([^>]+?)\s*>", line)
            if m:
                block = m.group(1)
                if block == "par_end":
                    break
                p.blocks.setdefault(block, {})
                continue
            if "=" in line and block is not None:
                k, v = line.split("=", 1)
                p.blocks[block][k.strip()] = v.strip()
        return p

    # ---- par_cmdline (par.c:194): block/name=value -------------------------------------------------
    def par_cmdline(self, args) -> None:
        for a in args:
            m = re.match(r"([^/=]+)/([^=]+)=(.*)", a)
            if not m:
                raise ParError(f"[par_cmdline]: cannot parse '{a}' (expected block/name=value)")
            self.blocks.setdefault(m.group(1).strip(), {})[m.group(2).strip()] = m.group(3).strip()

    def par_exist(self, block: str, name: str) -> bool:
        return block in self.blocks and name in self.blocks[block]

    def _get(self, block: str, name: str) -> str:
        if block not in self.blocks:
            raise ParError(f"[par_get]: Block name '{block}' not found")
        if name not in self.blocks[block]:
            raise ParError(f"[par_get]: Par name '{name}' not found in Block '{block}'")
        return self.blocks[block][name]

    def par_gets(self, block, name) -> str:
        return self._get(block, name)

    def par_geti(self, block, name) -> int:
        # par.c uses atoi(): leading integer part only ("100000000." -> 100000000)
        m = re.match(r"\s*[-+]?\d+", self._get(block, name))
        return int(m.group(0)) if m else 0

    def par_getd(self, block, name) -> float:
        return float(self._get(block, name))

    def par_geti_def(self, block, name, default: int) -> int:
        return self.par_geti(block, name) if self.par_exist(block, name) else default

    def par_getd_def(self, block, name, default: float) -> float:
        return self.par_getd(block, name) if self.par_exist(block, name) else default

    def par_dump(self) -> str:
        out = []
        for b, kv in self.blocks.items():
            out.append(f"<{b}>")
            out += [f"{k} = {v}" for k, v in kv.items()]
            out.append("")
        return "\n".join(out)
