package PhyloProf::Algorithm::ppp_b200;

# Drop-in for PhyloProf::Algorithm::ppp (lib/PhyloProf/Algorithm/ppp.pm:58-245): same constructor arguments
# (named -prof1,-prof2,[-prob],[-rank],[-taxonomy],[-dup] or positional, Root.pm:5-17), same 5-element return
# list of get_match_score(), scored on the B200 through PhyloProf::B200.
#
# A single pair is embedded exactly by taking the TARGET's own de-duplicated list as the genome order: the target
# plane is then a run of ones, the query planes are permuted to match, and the raw list position (duplicates and
# foreign taxa included, ppp.pm:169,231) is restored from the de-duplicated depth on return.  With -dup the pair goes
# through the ranked-list entry instead (PhyloProf::B200::score_ranked, which keeps the list's first repeated entry).
use strict;
use warnings;
use Carp;
use PhyloProf::B200;

my $GPU;    # one context per process, created lazily (after any fork)

sub new {
    my $class = shift;
    my %a;
    if ( @_ && defined $_[0] && !ref( $_[0] ) && $_[0] =~ /^-/ ) {
        my %in = @_;
        $a{ lc( substr( $_, 1 ) ) } = $in{$_} for keys %in;
    } else {
        @a{qw(prof1 prof2 prob rank taxonomy dup)} = @_;
    }
    croak "Parameter is undefined\n" unless $a{prof1} || $a{prof2};
    croak "Parameter mismatch\n"
        unless ( ref $a{prof1} && $a{prof1}->isa("PhyloProf::Profile") ) || ( ref $a{prof2} && $a{prof2}->isa("PhyloProf::Profile") );
    my $self = bless { _prof1 => $a{prof1}, _prof2 => $a{prof2} }, ref($class) || $class;
    $self->{_prob} = $a{prob}     if $a{prob};
    $self->{_rank} = $a{rank}     if $a{rank};
    $self->{_dbh}  = $a{taxonomy} if $a{taxonomy};
    $self->{_dup}  = 1            if $a{dup};
    return $self;
}

sub _taxon {    # ppp.pm:328-370: identity unless -rank; collapse, fall back to genus, then to the raw id
    my ( $self, $t ) = @_;
    return $t unless $self->{_rank};
    croak "Taxonomy does not exist\n" unless exists $self->{_dbh};    # ppp.pm:345-349 (callers may also poke $alg->{_dbh}, bin/ppp.pl:572)
    return $self->{_dbh}->collapse_taxon( $t, $self->{_rank} ) || $self->{_dbh}->collapse_taxon( $t, "genus" ) || $t;
}

sub get_match_score {
    my $self = shift;
    my $q = $self->{_prof1}->get_hash();
    my $p = exists $self->{_prob} ? $self->{_prob} : $self->{_prof1}->get_yes() / $self->{_prof1}->get_total();
    my $t = $self->{_prof2}->get_hash();
    my @list = ref($t) eq 'HASH' ? ( sort { $t->{$a} <=> $t->{$b} } keys %$t ) : @{ $t || [] };
    if ( $self->{_dup} ) {
        # -dup (ppp.pm:189-200, 222-224): the list keeps its own order and its first repeated entry; ranked-list path
        my @taxa = map { my $x = $self->_taxon($_); die "Taxon not found" unless $x; $x } @list;
        return ( 1, 0, 0, 0, $p ) unless @taxa && %$q;
        my @universe = keys %$q;
        croak "query profile of " . scalar(@universe) . " taxa exceeds 8192" if @universe > 8192;
        $GPU ||= PhyloProf::B200->new( -device => $ENV{PPP_B200_DEVICE} || 0 );
        my $hits = $GPU->score_ranked( \@universe, [$q], [ \@taxa ], -probs => [$p], -dup => 1 );
        my ( undef, undef, $score, $yes, $number, $depth ) = @{ $hits->[0] };
        return ( $score, $yes, $number, $depth, $p );
    }
    my ( %seen, @order, @rawpos );
    for my $i ( 0 .. $#list ) {
        my $taxon = $self->_taxon( $list[$i] );
        die "Taxon not found" unless $taxon;    # ppp.pm:179
        next if $seen{$taxon}++;                # ppp.pm:189-192
        push @order,  $taxon;
        push @rawpos, $i + 1;
    }
    return ( 1, 0, 0, 0, $p ) unless @order;
    croak "target list of " . scalar(@order) . " distinct taxa exceeds 8192" if @order > 8192;
    $GPU ||= PhyloProf::B200->new( -device => $ENV{PPP_B200_DEVICE} || 0 );
    my %qh = map { $_ => $q->{$_} } grep { $seen{$_} } keys %$q;    # taxa outside the target never take part
    my $hits = $GPU->score_bits( \@order, [ \%qh ], [ \@order ], -probs => [$p] );
    my ( undef, undef, $score, $yes, $number, $depth ) = @{ $hits->[0] };
    $depth = $rawpos[ $depth - 1 ] if $depth > 0;
    return ( $score, $yes, $number, $depth, $p );
}

1;
