"""Host-side mirrors of dn_splatter/utils helpers that sit on the hot path."""
