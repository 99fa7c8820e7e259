"""Model side of the drop-in surface: CFM (sampler) and the DiT backbone (reference lemas_tts/model/)."""
