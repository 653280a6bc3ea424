"""Mirror of the reference's ``libs`` package (pvlt, vl_heads, vl_scores)."""
