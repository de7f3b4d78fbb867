"""Host-side mirror of the reference's `lib/` solve-phase modules."""
