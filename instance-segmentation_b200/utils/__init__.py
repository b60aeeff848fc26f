"""Drop-in mirrors of the reference's utils/ modules that lie on the decode hot path."""
