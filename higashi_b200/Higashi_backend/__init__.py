"""Mirror of the reference package `higashi.Higashi_backend` (hot-path subset)."""
