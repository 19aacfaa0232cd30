from openpvsg_b200.relation_head import concatenate_sub_obj  # noqa: F401
