"""B200-native two-view augmentation + contrastive loss."""
