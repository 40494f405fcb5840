function [s, R, t, time_taken] = teaser_solve(src, dst, varargin)
%TEASER_SOLVE  Name/value wrapper around teaser_solve_mex (same interface as the reference's matlab/teaser_solve.m).
%   [s, R, t, time_taken] = teaser_solve(src, dst, 'NoiseBound', 0.01, 'Cbar2', 1, 'EstimateScaling', true, ...
%       'RotationEstimationAlgorithm', 0, 'RotationGNCFactor', 1.4, 'RotationMaxIterations', 100, ...
%       'RotationCostThreshold', 1e-6, 'InlierSelectionAlgorithm', 0, 'KcoreHeuristicThreshold', 0.5)
%   time_taken is returned in seconds.
p = inputParser;
addParameter(p, 'Cbar2', 1);
addParameter(p, 'NoiseBound', 0.01);
addParameter(p, 'EstimateScaling', true);
addParameter(p, 'RotationEstimationAlgorithm', 0);
addParameter(p, 'RotationGNCFactor', 1.4);
addParameter(p, 'RotationMaxIterations', 100);
addParameter(p, 'RotationCostThreshold', 1e-6);
addParameter(p, 'InlierSelectionAlgorithm', 0);
addParameter(p, 'KcoreHeuristicThreshold', 0.5);
parse(p, varargin{:});
q = p.Results;
[s, R, t, ms] = teaser_solve_mex(double(src), double(dst), double(q.Cbar2), double(q.NoiseBound), ...
    logical(q.EstimateScaling), double(q.RotationEstimationAlgorithm), double(q.RotationGNCFactor), ...
    double(q.RotationMaxIterations), double(q.RotationCostThreshold), double(q.InlierSelectionAlgorithm), ...
    double(q.KcoreHeuristicThreshold));
time_taken = ms / 1000;
end
