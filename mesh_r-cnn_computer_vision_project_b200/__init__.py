# Sources of the ``meshrcnn_b200`` package (see ../meshrcnn_b200/__init__.py for the import alias).
