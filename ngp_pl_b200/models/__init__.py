"""Host-side mirror of the reference's `models` package (custom_functions, networks, rendering)."""
