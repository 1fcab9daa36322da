"""B200-native fiducial detection for the MAMRI pose-estimation pipeline."""
