"""Response model of the colour service.

Inside the reference application the service returns the app's own pydantic
model (app/api/v1/models/analyze.py:157-204); when that module is importable it
is used as-is, otherwise an equivalent model with the same fields, validation
pattern and `from_dict` helper is defined here."""
from __future__ import annotations

import re
from typing import Any, Dict, List, Optional

try:  # running inside the reference application
    from app.api.v1.models.analyze import ColorFeatures  # type: ignore  # noqa: F401
except Exception:  # standalone
    from pydantic import BaseModel, Field, field_validator

    _HEX = r"^#(?:[0-9a-fA-F]{3}){1,2}$"

    class ColorFeatures(BaseModel):
        """Model for color extraction results."""
        primary: Optional[str] = Field(None, pattern=_HEX)
        background: Optional[str] = Field(None, pattern=_HEX)
        accent: List[str] = Field(default_factory=list)
        metadata: Dict[str, Any] = Field(default_factory=dict)

        @field_validator("accent")
        @classmethod
        def _validate_accent(cls, v):
            for c in v:
                if not re.match(_HEX, c):
                    raise ValueError(f"Invalid hex color code: {c}")
            return v

        @classmethod
        def from_dict(cls, data: Dict[str, Any]) -> "ColorFeatures":
            md = data.get("metadata", {})
            return cls(primary=data.get("primary"), background=data.get("background"), accent=data.get("accent", []),
                       metadata={"success": md.get("success", True), "timestamp": md.get("timestamp", 0.0),
                                 "processing_time": md.get("processing_time", 0.0)})
