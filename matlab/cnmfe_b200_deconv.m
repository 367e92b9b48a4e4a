function s = cnmfe_b200_deconv(opts)
%% options.deconv_options -> plain struct with the field names of deconvolveCa's parser (OASIS_matlab/deconvolveCa.m:167-356).
% deconvolveCa accepts a struct, or name/value pairs in a cell; [] = its defaults (ar1, constrained).
if isempty(opts)
    s = struct();
elseif isstruct(opts)
    s = opts;
elseif iscell(opts)
    s = struct();
    k = 1;
    while k <= numel(opts)
        v = opts{k};
        if isstruct(v)                           % an option struct among the arguments (deconvolveCa.m:188-200)
            f = fieldnames(v);
            for i = 1:numel(f); s.(f{i}) = v.(f{i}); end
            k = k + 1;
        elseif ischar(v)
            switch lower(v)
                case {'ar1', 'ar2', 'exp2', 'kernel'}; s.type = lower(v); k = k + 1;
                case {'foopsi', 'constrained', 'thresholded', 'mcmc'}; s.method = lower(v); k = k + 1;
                case 'optimize_b';    s.optimize_b = true; k = k + 1;
                case 'optimize_pars'; s.optimize_pars = true; k = k + 1;
                case 'optimize_smin'; s.optimize_smin = true; k = k + 1;
                otherwise
                    s.(lower(v)) = opts{k+1};    % 'pars', 'sn', 'b', 'lambda', 'smin', 'maxiter', 'window', 'shift', 'tau_range', 'max_tau', ...
                    k = k + 2;
            end
        else
            k = k + 1;
        end
    end
else
    error('cnmfe:b200', 'deconv_options must be a struct, a cell of deconvolveCa arguments, or empty');
end
if isfield(s, 'maxiter') && ~isfield(s, 'maxIter'); s.maxIter = s.maxiter; end
end
