"""Host-side mirror of the reference's ``nerf`` package for the render hot path only
(``nerf/renderer.py`` + ``nerf/network.py``); trainer, data providers and GUI are out of scope."""
