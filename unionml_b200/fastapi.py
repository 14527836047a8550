"""FastAPI binding: ``POST /predict``, ``GET /health``, ``GET /`` on the device path.

Same routes, request body and error behaviour as ``/root/reference/unionml/fastapi.py:15-70``:

* startup loads the model from ``UNIONML_MODEL_PATH`` when no artifact is set (ref. 22-34); the CUDA engine is *not*
  touched here - it is bound lazily by the first predict in the worker process (uvicorn ``--workers`` forks);
* ``/predict`` takes ``{"features": [...]}`` or ``{"inputs": {...reader kwargs}}``; neither -> HTTP 500 (ref. 55-56);
* ``get_features`` runs before ``Model.predict`` and again inside it (ref. 61 + ``model.py:740``) - kept as written;
* ``/health`` -> 500 until a model artifact exists (ref. 66-70).

Remote (Flyte cluster) artifact resolution is out of scope; ``remote=True`` raises.
"""

import os
from http import HTTPStatus
from typing import Any, Dict, List, Optional

from fastapi import Body, FastAPI, HTTPException
from fastapi.responses import HTMLResponse

from unionml_b200.model import Model, ModelArtifact


def serving_app(
    model: Model,
    app: FastAPI,
    remote: bool = False,
    app_version: Optional[str] = None,
    model_version: str = "latest",
):
    if remote:
        raise NotImplementedError("remote=True (Flyte-backed artifacts) is out of scope for unionml_b200")

    def setup_model():
        model_path = os.getenv("UNIONML_MODEL_PATH")
        if model.artifact is None:
            if model_path is None:
                raise ValueError(
                    "Model artifact path not specified. Make sure to specify the unionml serve --model-path in "
                    "the option when starting the unionml prediction service in local mode."
                )
            model.artifact = ModelArtifact(model.load(model_path))

    app.router.on_startup.append(setup_model)

    @app.get("/", response_class=HTMLResponse)
    def root():
        return "<html><head><title>unionml</title></head><body><h1>unionml</h1><p>B200 predict engine</p></body></html>"

    @app.post("/predict")
    async def predict(
        inputs: Optional[Dict[str, Any]] = Body(None),
        features: Optional[List[Any]] = Body(None),
    ):
        if inputs is None and features is None:
            raise HTTPException(status_code=500, detail="inputs or features must be supplied.")
        workflow_inputs: Dict[str, Any] = {}
        if model._dataset.dataset_datatype is not None:
            features = model._dataset.get_features(features)
        workflow_inputs.update(inputs if inputs else {"features": features})
        return model.predict(**workflow_inputs)

    @app.get("/health")
    async def health():
        if model.artifact is None:
            raise HTTPException(status_code=500, detail="Model artifact not found.")
        return {"message": HTTPStatus.OK.phrase, "status": HTTPStatus.OK}

    return app
