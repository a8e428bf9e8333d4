"""Import surface mirroring the reference package layout (pioneer/envs/{bullet,pioneer})."""
