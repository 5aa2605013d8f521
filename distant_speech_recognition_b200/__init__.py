"""distant_speech_recognition_b200 — B200-native subband beamforming pipe (drop-in for btk2.0's hot path)."""
