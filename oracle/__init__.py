"""Test-only CPU oracle for the UnCRtainTS hot path (see uncrtaints_oracle.py header).
Never imported by the product package ``uncrtaints_b200``."""
