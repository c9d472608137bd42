"""The reference's template table, restated in closed form.

`TEMPLATE_CONFIGS[(num_bits, template_id)]` is what `flute/__init__.py:53-69` loads from
`flute/data/qgemm_kernel_raw_generated_configs.pth`; `flute/codegen_utils.py:89-160` is the
enumeration that file was written from.  The new engine is not template-enumerated -- the only
field it consumes is "TileP" (which packing a stored checkpoint uses, `flute/utils.py:302-309`)
-- but callers index the table (`integrations/base.py:157`, `huggingface.py:190-209`), so it is
kept, key for key.  tests/test_templates.py checks it against the reference's file (golden).
"""
from __future__ import annotations

from typing import Dict, Tuple

_SMS_MULTIPLE = (1, 2, 4)
_TILES = ((256, 32, 64, 64), (256, 32, 64, 32), (128, 16, 64, 32))   # Threads, TileM, TileK, TileP
_STAGES = (2, 3, 4, 5)
_QMODES = ("kVectorized   ", "kVectorized_32", "kVectorized_16", "kVectorized_8 ")


def build_template_configs() -> Dict[Tuple[int, int], Dict]:
    configs: Dict[Tuple[int, int], Dict] = {}
    for num_bits in (4, 3, 2):
        index = 0
        for sms in _SMS_MULTIPLE:
            for threads, tile_m, tile_k, tile_p in _TILES:
                for stages in _STAGES:
                    for qmode in (_QMODES if num_bits == 4 else _QMODES[:1]):
                        configs[(num_bits, index)] = {
                            "SMs_Multiple": sms,
                            "Threads": threads,
                            "TileM": tile_m,
                            "TileK": tile_k,
                            "TileP": tile_p,
                            "Stages": stages,
                            "QuantMapMode": qmode,
                            "AccumulationMode": "kMixed",
                            "DecompositionMode": "kStreamK",
                            "G2STiledCopySizeS": 2,
                            "MmaPrmK": 1,
                        }
                        index += 1
    return configs


TEMPLATE_CONFIGS = build_template_configs()


def tile_p_of(num_bits: int, template_id: int) -> int:
    try:
        return TEMPLATE_CONFIGS[(num_bits, template_id)]["TileP"]
    except KeyError:
        raise RuntimeError(f"Unsupported template_id {template_id} for num_bits {num_bits}") from None


def default_template_id(num_bits: int) -> int:
    """The id the engine hands out when it packs: first TileP == 32 entry
    (every shipped tuned 4-bit config and all 3-bit packing use TileP = 32)."""
    for (b, i), cfg in sorted(TEMPLATE_CONFIGS.items()):
        if b == num_bits and cfg["TileP"] == 32:
            return i
    raise RuntimeError(num_bits)
