"""B200-native FINCH clustering / nearest-neighbour retrieval hot path of SLIC
(rvl-lab-utoronto/video_similarity_search), behind the reference's own Python signatures.

    from video_similarity_search_b200.clustering.finch import FINCH
    from video_similarity_search_b200.clustering.cluster_masks import fit_cluster, positive_mask, negative_mask
    from video_similarity_search_b200.evaluate import get_distance_matrix, get_closest_data_mat, get_topk_acc
    from video_similarity_search_b200.iic_retrieve_clips import topk_retrieval

All computation is hand-written sm_100a CUDA in libslic_b200.so (C ABI: include/slic_b200.h).
"""
__version__ = "0.1.0"
